T=${1:-s2i}
# launch list of the headline step (shares), full capture of the two GEMM-class kernels, PGD launch list
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 60 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --no-cpu-baseline --no-extra --steps 2 --warmup 3 > gpurun_out/${T}_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'fc_fused_kernel|tc_gemm_kernel' -s 18 -c 2 -o gpurun_out/${T}_f16x3_full python bench.py --no-cpu-baseline --no-extra --steps 1 --warmup 3 > gpurun_out/${T}_full.log 2>&1
RBNN_PGD_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file gpurun_out/${T}_pgd_launches.csv python scratch/pgd_probe.py 6 > gpurun_out/${T}_pgd_ncu.log 2>&1
python scratch/pgd_probe.py 20 > gpurun_out/${T}_pgd.log 2>&1
RBNN_PGD_GRAPH=0 python scratch/pgd_probe.py 20 >> gpurun_out/${T}_pgd.log 2>&1
cat gpurun_out/${T}_pgd.log
