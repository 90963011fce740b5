T=${1:-s3o}
timeout 900 python -m pytest tests -m gpu -x -q -k "conv" > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -4 gpurun_out/${T}_pytest.log
python scratch/conv_probe.py f16x3 > gpurun_out/${T}_conv_probe.json 2> gpurun_out/${T}_conv_probe.err; cat gpurun_out/${T}_conv_probe.json; tail -2 gpurun_out/${T}_conv_probe.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_conv_launches.csv python scratch/conv_probe.py f16x3 2 > gpurun_out/${T}_conv_ncu.log 2>&1
python profiles/extract_ncu.py --launches gpurun_out/${T}_conv_launches.csv 2>/dev/null | head -10
