T=${1:-s3s}
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -5 gpurun_out/${T}_pytest.log
timeout 600 python bench.py --no-cpu-baseline --no-extra > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${T}_bench.json').read().strip().splitlines()[-1])
r=d['roofline']
print(d['ms_per_step'], r['kernel'][:16], r['avg_launch_ms'], r['frac'], r['other_gemm_class_ms'], d['e2e']['ms_per_step'], d.get('float_inputs',{}).get('ms_per_step'))
PY
python scratch/conv_probe.py f16x3 > gpurun_out/${T}_conv_probe.json 2> gpurun_out/${T}_conv_probe.err; cat gpurun_out/${T}_conv_probe.json
python scratch/pgd_probe.py 20 2>&1 | tail -1
python scratch/fc2_probe.py tf32x3 2>&1 | tail -1
