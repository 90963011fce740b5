T=${1:-s2d}
RBNN_FUSED_DEBUG=8 python bench.py --no-cpu-baseline --no-extra --steps 1 --warmup 1 > gpurun_out/${T}_timers.log 2>&1
grep "fused cta0" gpurun_out/${T}_timers.log | tail -6
RBNN_FUSED_KBB=64 python bench.py --no-cpu-baseline --no-extra > gpurun_out/${T}_bench_kbb64.json 2> gpurun_out/${T}_bench_kbb64.err
ncu --set full --clock-control none --import-source on -k regex:'fc_fused_kernel' -s 9 -c 1 -o gpurun_out/${T}_fused python bench.py --no-cpu-baseline --no-extra --steps 1 --warmup 3 > gpurun_out/${T}_full.log 2>&1
for f in gpurun_out/${T}_bench*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['roofline']['other_gemm_class_ms'], d['roofline']['kernel'][:20], d.get('clocks'))
"; done
