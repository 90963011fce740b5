# ncu full capture (with source) of the fused kernel + backward GEMM on the current tree
ncu --set full --clock-control none --import-source on -k regex:'fc_fused_kernel|tc_gemm_kernel' -s 18 -c 2 -o gpurun_out/s2b_fused python bench.py --no-cpu-baseline --no-extra --steps 1 --warmup 3 > gpurun_out/s2b_full.log 2>&1
echo done
