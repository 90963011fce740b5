T=${1:-s3q}
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 80 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --no-cpu-baseline --no-extra --steps 2 --warmup 3 > gpurun_out/${T}_ncu_bench.log 2>&1
python profiles/extract_ncu.py --launches gpurun_out/${T}_launches.csv 2>/dev/null | head -14
ncu --set full --clock-control none --import-source on -k regex:'fc_fused_kernel|tc_gemm_kernel' -s 12 -c 2 -o gpurun_out/${T}_f16x3_full python bench.py --no-cpu-baseline --no-extra --steps 1 --warmup 3 > gpurun_out/${T}_full.log 2>&1
tail -2 gpurun_out/${T}_full.log
ncu --set full --clock-control none --import-source on -k regex:'conv1_pool_fwd|p1_split|tc_gemm_kernel|conv2_refine|pool2_logits|pool2_bwd|col2im|conv1_bwd_sum' -s 9 -c 9 -o gpurun_out/${T}_conv_full python scratch/conv_probe.py f16x3 1 > gpurun_out/${T}_conv_full.log 2>&1
tail -2 gpurun_out/${T}_conv_full.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_conv_launches.csv python scratch/conv_probe.py f16x3 2 > gpurun_out/${T}_conv_ncu.log 2>&1
