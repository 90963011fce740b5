N=${1:-8}; T=${2:-s3k}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --no-cpu-baseline > gpurun_out/${T}_bench_${N}gpu.json 2> gpurun_out/${T}_bench_${N}gpu.err
tail -2 gpurun_out/${T}_bench_${N}gpu.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${T}_bench_${N}gpu.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['ms_per_step'], d['e2e']['ms_per_step'], d.get('pgd',{}).get('ms'), d.get('halfmoons',{}).get('ms'), d.get('clocks'))
PY
if [ "$N" = "2" ]; then timeout 600 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -3; fi
