T=${1:-s4a}
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log; tail -3 gpurun_out/${T}_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; python -c "
import json;d=json.loads(open('gpurun_out/${T}_bench_ref.json').read().strip().splitlines()[-1]);print('ref', d['value'], d['cpu_baseline']['kind'], d['cpu_baseline']['cores'])"
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; python -c "
import json;d=json.loads(open('gpurun_out/${T}_bench.json').read().strip().splitlines()[-1]);r=d['roofline'];print('ours', d['ms_per_step'], d['e2e']['ms_per_step'], r['frac'], r['avg_launch_ms'], d['cpu_baseline']['value'], d['gpu_launches'], d['clocks'])"
