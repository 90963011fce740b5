T=${1:-s4d}
timeout 1200 python -m pytest tests -m gpu -x -q -k "two_pass or headline or tcgen05 or golden or pgd or saturated or keep or graph or ensemble" > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -8 gpurun_out/${T}_pytest.log
for F in 1 0 1 0; do
RBNN_FUSED_FRAG=$F timeout 300 python bench.py --no-cpu-baseline --no-extra > gpurun_out/${T}_bench_f$F.json 2> gpurun_out/${T}_bench_f$F.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${T}_bench_f$F.json').read().strip().splitlines()[-1])
r=d['roofline']
fwd = r['other_gemm_class_ms'] if r['kernel'].startswith('tc_gemm') else r['avg_launch_ms']*r['launches']
print('frag=$F step', round(d['ms_per_step'],2), ' fwd per chunk', round(fwd/r['launches'],3), ' float-input step', round(d.get('float_inputs',{}).get('ms_per_step',0),2))
PY
done
