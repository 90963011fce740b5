T=${1:-s2c}
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -3 gpurun_out/${T}_pytest.log
python bench.py --no-cpu-baseline --no-extra > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
RBNN_FUSED_DEBUG=1 python bench.py --no-cpu-baseline --no-extra > gpurun_out/${T}_bench_nostore.json 2> gpurun_out/${T}_bench_nostore.err
RBNN_FUSED_DEBUG=3 python bench.py --no-cpu-baseline --no-extra > gpurun_out/${T}_bench_nopass2.json 2> gpurun_out/${T}_bench_nopass2.err
RBNN_FUSED_DEBUG=7 python bench.py --no-cpu-baseline --no-extra > gpurun_out/${T}_bench_noepi.json 2> gpurun_out/${T}_bench_noepi.err
for f in gpurun_out/${T}_bench*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['roofline']['other_gemm_class_ms'], d['roofline']['kernel'][:20], d.get('clocks'))
"; done
