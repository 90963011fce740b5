T=${1:-s3j}
timeout 1500 python bench.py > gpurun_out/${T}_bench_full.json 2> gpurun_out/${T}_bench_full.err
tail -3 gpurun_out/${T}_bench_full.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err
cat gpurun_out/${T}_bench_ref.json | head -c 1500
python - <<PY
import json
d=json.loads(open('gpurun_out/${T}_bench_full.json').read().strip().splitlines()[-1])
r=d['roofline']
print(d['ms_per_step'], r['kernel'][:16], r['avg_launch_ms'], r['frac'], r['other_gemm_class_ms'], d['e2e']['ms_per_step'], d.get('float_inputs',{}).get('ms_per_step'))
for k in ('pgd','fgsm','halfmoons','conv_cfg4','sampler','library_gpu_baseline','cpu_baseline'):
    print(k, json.dumps(d.get(k))[:400])
PY
