T=${1:-s3e}
timeout 1200 python -m pytest tests -m gpu -x -q -k "two_pass or headline or tcgen05 or golden" > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -4 gpurun_out/${T}_pytest.log
timeout 600 python bench.py --no-cpu-baseline --no-extra > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python - <<PY
import json
for f in ('gpurun_out/${T}_bench.json',):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d['roofline']
        print(f, d['ms_per_step'], r['kernel'][:16], r['avg_launch_ms'], r['frac'], r['other_gemm_class_ms'], r.get('forward_passes'), d['e2e']['ms_per_step'], d.get('float_inputs',{}).get('ms_per_step'), d.get('clocks'))
    except Exception as e:
        print(f, 'failed', e)
PY
tail -3 gpurun_out/${T}_bench.err
RBNN_FUSED_DEBUG=8 python bench.py --no-cpu-baseline --no-extra --steps 1 --warmup 3 --samples 334 2>&1 | grep "fused cta0" | tail -2 > gpurun_out/${T}_timers.log
cat gpurun_out/${T}_timers.log
