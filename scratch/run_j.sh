T=${1:-s2j}
timeout 600 python -m pytest tests -m gpu -x -q -k "sampler or philox or pgd or golden_svi" > gpurun_out/${T}_pytest.log 2>&1
tail -3 gpurun_out/${T}_pytest.log
RBNN_PGD_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file gpurun_out/${T}_pgd_launches.csv python scratch/pgd_probe.py 6 > gpurun_out/${T}_pgd_ncu.log 2>&1
python scratch/pgd_probe.py 20 > gpurun_out/${T}_pgd.log 2>&1
RBNN_PGD_GRAPH=0 python scratch/pgd_probe.py 20 >> gpurun_out/${T}_pgd.log 2>&1
cat gpurun_out/${T}_pgd.log
