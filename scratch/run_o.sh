T=${1:-s3b}
python scratch/conv_cfg4_oracle.py > gpurun_out/${T}_conv_cfg4_oracle.json 2> gpurun_out/${T}_conv_cfg4_oracle.err
cat gpurun_out/${T}_conv_cfg4_oracle.json; tail -3 gpurun_out/${T}_conv_cfg4_oracle.err
timeout 900 python -m pytest tests -m gpu -x -q -k conv > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -5 gpurun_out/${T}_pytest.log
python scratch/conv_probe.py f16x3 > gpurun_out/${T}_conv_probe.json 2> gpurun_out/${T}_conv_probe.err
cat gpurun_out/${T}_conv_probe.json
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_conv_launches.csv python scratch/conv_probe.py f16x3 2 > gpurun_out/${T}_conv_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'conv1_pool_fwd|p1_split|tc_gemm_kernel|conv2_refine|pool2_logits|pool2_bwd|col2im|conv1_bwd_sum' -s 9 -c 9 -o gpurun_out/${T}_conv_full python scratch/conv_probe.py f16x3 1 > gpurun_out/${T}_conv_full.log 2>&1
tail -3 gpurun_out/${T}_conv_full.log
