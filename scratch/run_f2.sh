for P in 1 2; do echo PAIRS=$P; RBNN_POOL_PAIRS=$P python scratch/conv_probe.py f16x3 2>&1 | tail -1 | cut -c60-200
RBNN_POOL_PAIRS=$P ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s4b_l$P.csv python scratch/conv_probe.py f16x3 1 > /dev/null 2>&1
python profiles/extract_ncu.py --launches gpurun_out/s4b_l$P.csv 2>/dev/null | grep -E "pool2" ; done
RBNN_POOL_PAIRS=2 python -m pytest tests -m gpu -x -q -k "conv" 2>&1 | tail -2
