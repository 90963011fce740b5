T=${1:-s3n}
timeout 900 python -m pytest tests -m gpu -x -q -k "conv or ensemble or golden" > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -6 gpurun_out/${T}_pytest.log
python scratch/conv_probe.py f16x3 > gpurun_out/${T}_conv_probe.json 2> gpurun_out/${T}_conv_probe.err; cat gpurun_out/${T}_conv_probe.json; tail -2 gpurun_out/${T}_conv_probe.err
HIDDEN=1024 python scratch/conv_probe.py f16x3 > gpurun_out/${T}_conv_probe_1024.json 2>&1; cat gpurun_out/${T}_conv_probe_1024.json
python scratch/conv_probe.py fp32 > gpurun_out/${T}_conv_probe_fp32.json 2>&1; cat gpurun_out/${T}_conv_probe_fp32.json
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_conv_launches.csv python scratch/conv_probe.py f16x3 2 > gpurun_out/${T}_conv_ncu.log 2>&1
python profiles/extract_ncu.py --launches gpurun_out/${T}_conv_launches.csv 2>/dev/null | head -12
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "conv and not cfg4" > gpurun_out/${T}_memcheck.log 2>&1
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/${T}_memcheck.log | tail -3
