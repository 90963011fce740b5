T=${1:-s3g}
timeout 300 robustbnns_b200/csrc/build/tc_gemm_test > gpurun_out/${T}_harness.log 2>&1
grep -c " ok" gpurun_out/${T}_harness.log; grep -n "FAIL" gpurun_out/${T}_harness.log | head; grep -n "cfg4\|TFLOP" gpurun_out/${T}_harness.log | tail -8
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -8 gpurun_out/${T}_pytest.log
for P in 3 0; do
RBNN_CONV_PAIR=$P python scratch/conv_probe.py f16x3 > gpurun_out/${T}_conv_probe_p$P.json 2> gpurun_out/${T}_conv_probe_p$P.err
echo "pair=$P"; cat gpurun_out/${T}_conv_probe_p$P.json; tail -2 gpurun_out/${T}_conv_probe_p$P.err
done
python scratch/conv_probe.py tf32x3 > gpurun_out/${T}_conv_probe_tf32.json 2>&1; cat gpurun_out/${T}_conv_probe_tf32.json
python scratch/fc2_probe.py > gpurun_out/${T}_fc2_probe.json 2>&1; cat gpurun_out/${T}_fc2_probe.json
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_conv_launches.csv python scratch/conv_probe.py f16x3 2 > gpurun_out/${T}_conv_ncu.log 2>&1
python profiles/extract_ncu.py --launches gpurun_out/${T}_conv_launches.csv 2>/dev/null | head -12
timeout 600 python bench.py --no-cpu-baseline --no-extra > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${T}_bench.json').read().strip().splitlines()[-1])
r=d['roofline']
print(d['ms_per_step'], r['kernel'][:16], r['avg_launch_ms'], r['frac'], r['other_gemm_class_ms'], d['e2e']['ms_per_step'], d.get('float_inputs',{}).get('ms_per_step'))
PY
