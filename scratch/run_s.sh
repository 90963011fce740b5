T=${1:-s3f}
timeout 900 python -m pytest tests -m gpu -x -q -k conv > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -8 gpurun_out/${T}_pytest.log
for P in 3 0 1 2; do
RBNN_CONV_PAIR=$P python scratch/conv_probe.py f16x3 > gpurun_out/${T}_conv_probe_p$P.json 2> gpurun_out/${T}_conv_probe_p$P.err
echo "pair=$P"; cat gpurun_out/${T}_conv_probe_p$P.json; tail -2 gpurun_out/${T}_conv_probe_p$P.err
done
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_conv_launches.csv python scratch/conv_probe.py f16x3 2 > gpurun_out/${T}_conv_ncu.log 2>&1
python profiles/extract_ncu.py --launches gpurun_out/${T}_conv_launches.csv | head -12
