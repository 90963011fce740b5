T=${1:-s3a}
python scratch/conv_cfg4_oracle.py > gpurun_out/${T}_conv_cfg4_oracle.json 2> gpurun_out/${T}_conv_cfg4_oracle.err
cat gpurun_out/${T}_conv_cfg4_oracle.json
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -5 gpurun_out/${T}_pytest.log
python scratch/conv_probe.py > gpurun_out/${T}_conv_probe.json 2> gpurun_out/${T}_conv_probe.err
cat gpurun_out/${T}_conv_probe.json
