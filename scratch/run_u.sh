T=${1:-s3h}
python scratch/fc2_probe.py tf32x3 > gpurun_out/${T}_fc2_probe.json 2> gpurun_out/${T}_fc2_probe.err; cat gpurun_out/${T}_fc2_probe.json; tail -3 gpurun_out/${T}_fc2_probe.err
ORACLE=0 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_fc2_launches.csv python scratch/fc2_probe.py tf32x3 1 > gpurun_out/${T}_fc2_ncu.log 2>&1
python profiles/extract_ncu.py --launches gpurun_out/${T}_fc2_launches.csv 2>/dev/null | head -14
