T=${1:-s2e}
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -3 gpurun_out/${T}_pytest.log
RBNN_FUSED_DEBUG=8 python bench.py --no-cpu-baseline --no-extra --steps 1 --warmup 1 > gpurun_out/${T}_timers.log 2>&1
grep "fused cta0" gpurun_out/${T}_timers.log | tail -3
python bench.py --no-cpu-baseline --no-extra > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
RBNN_FUSED_DEBUG=1 python bench.py --no-cpu-baseline --no-extra > gpurun_out/${T}_bench_nostore.json 2> gpurun_out/${T}_bench_nostore.err
for f in gpurun_out/${T}_bench*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['roofline']['other_gemm_class_ms'], d['roofline']['kernel'][:20], d.get('clocks'))
"; done
