T=${1:-s2k}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -4 gpurun_out/${T}_pytest.log
python scratch/pgd_probe.py 20 > gpurun_out/${T}_pgd.log 2>&1
python scratch/pgd_probe.py 20 >> gpurun_out/${T}_pgd.log 2>&1
cat gpurun_out/${T}_pgd.log
RBNN_PGD_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file gpurun_out/${T}_pgd_launches.csv python scratch/pgd_probe.py 6 > gpurun_out/${T}_pgd_ncu.log 2>&1
timeout 600 python bench.py --no-cpu-baseline --no-extra > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${T}_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['roofline']['other_gemm_class_ms'], d['e2e']['ms_per_step'], d.get('clocks'))
PY
