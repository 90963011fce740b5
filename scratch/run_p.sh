T=${1:-s3c}
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -6 gpurun_out/${T}_pytest.log
timeout 600 python bench.py --no-cpu-baseline --no-extra > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
RBNN_XGRID=0 timeout 600 python bench.py --no-cpu-baseline --no-extra > gpurun_out/${T}_bench_xgrid0.json 2> gpurun_out/${T}_bench_xgrid0.err
python - <<PY
import json
for f in ('gpurun_out/${T}_bench.json','gpurun_out/${T}_bench_xgrid0.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d['roofline']
        print(f, d['ms_per_step'], r['kernel'][:16], r['avg_launch_ms'], r['frac'], r['other_gemm_class_ms'], r.get('forward_passes'), d['e2e']['ms_per_step'], d.get('float_inputs',{}).get('ms_per_step'), d.get('clocks'))
    except Exception as e:
        print(f, 'failed', e)
PY
tail -5 gpurun_out/${T}_bench.err
