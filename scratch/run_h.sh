T=${1:-s2h}
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/${T}_pytest_multi.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${T}_pytest_multi.log
tail -4 gpurun_out/${T}_pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/${T}_bench_2gpu.json 2> gpurun_out/${T}_bench_2gpu.err
tail -3 gpurun_out/${T}_bench_2gpu.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${T}_bench_2gpu.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['e2e']['ms_per_step'], d.get('clocks'))
print('pgd', d.get('pgd')); print('halfmoons', {k:v for k,v in d.get('halfmoons',{}).items() if k in ('value','ms','e2e','n_gpus')})
PY
