T=${1:-s2f}
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -3 gpurun_out/${T}_pytest.log
RBNN_FUSED_PAIR=1 timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest_pair.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${T}_pytest_pair.log
tail -3 gpurun_out/${T}_pytest_pair.log
RBNN_FUSED_DEBUG=8 timeout 300 python bench.py --no-cpu-baseline --no-extra --steps 1 --warmup 1 > gpurun_out/${T}_timers.log 2>&1
grep "fused cta0" gpurun_out/${T}_timers.log | tail -3
timeout 300 python bench.py --no-cpu-baseline --no-extra > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
RBNN_FUSED_PAIR=0 timeout 300 python bench.py --no-cpu-baseline --no-extra > gpurun_out/${T}_bench_single.json 2> gpurun_out/${T}_bench_single.err
RBNN_FUSED_DEBUG=7 timeout 300 python bench.py --no-cpu-baseline --no-extra > gpurun_out/${T}_bench_noepi.json 2> gpurun_out/${T}_bench_noepi.err
for f in gpurun_out/${T}_bench*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['roofline']['other_gemm_class_ms'], d['roofline']['kernel'][:20], d.get('clocks'))
"; done
