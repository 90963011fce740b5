T=${1:-s3l}
for D in 0 32 64 1; do
RBNN_FUSED_DEBUG=$D timeout 300 python bench.py --no-cpu-baseline --no-extra --steps 2 --warmup 3 > gpurun_out/${T}_dbg$D.json 2> gpurun_out/${T}_dbg$D.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${T}_dbg$D.json').read().strip().splitlines()[-1])
    r=d['roofline']
    fwd = r['other_gemm_class_ms'] if r['kernel'].startswith('tc_gemm') else r['avg_launch_ms']*r['launches']
    print('debug=$D  step', round(d['ms_per_step'],2), ' fwd per chunk', round(fwd/r['launches'],3), ' float-input step', round(d.get('float_inputs',{}).get('ms_per_step',0),2))
except Exception as e:
    print('debug=$D failed', e)
PY
done
