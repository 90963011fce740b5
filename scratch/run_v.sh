T=${1:-s3i}
timeout 900 python -m pytest tests -m gpu -x -q -k "fc2 or tcgen05 or moons or best_precision" > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -5 gpurun_out/${T}_pytest.log
python scratch/fc2_probe.py tf32x3 > gpurun_out/${T}_fc2_probe.json 2> gpurun_out/${T}_fc2_probe.err; cat gpurun_out/${T}_fc2_probe.json; tail -3 gpurun_out/${T}_fc2_probe.err
ORACLE=0 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_fc2_launches.csv python scratch/fc2_probe.py tf32x3 1 > gpurun_out/${T}_fc2_ncu.log 2>&1
python profiles/extract_ncu.py --launches gpurun_out/${T}_fc2_launches.csv 2>/dev/null | head -10
