for i in 1 2 3; do
for T in old new; do
if [ $T = old ]; then cd scratch/old_tree; else cd /root/repo; fi
python bench.py --no-cpu-baseline --no-extra 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);r=d['roofline'];print('$T', round(d['ms_per_step'],2), r['kernel'][:8], round(r['avg_launch_ms'],3), 'other', round(r['other_gemm_class_ms']/9,3))"
cd /root/repo
done; done
