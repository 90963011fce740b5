T=${1:-s2g}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -5 gpurun_out/${T}_pytest.log
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
tail -5 gpurun_out/${T}_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${T}_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['e2e']['ms_per_step'], d.get('clocks'))
print('pgd', d.get('pgd')); print('halfmoons', d.get('halfmoons')); print('cpu', d.get('cpu_baseline'))
PY
