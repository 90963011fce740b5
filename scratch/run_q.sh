T=${1:-s3d}
python scratch/conv_cfg4_oracle.py > gpurun_out/${T}_conv_cfg4_oracle.json 2> gpurun_out/${T}_conv_cfg4_oracle.err
python - <<PY
import json
d=json.load(open('gpurun_out/${T}_conv_cfg4_oracle.json'))
for k in ('fp32','tf32x3','f16x3'): print(k, d[k]['vs_fp64'])
PY
tail -3 gpurun_out/${T}_conv_cfg4_oracle.err
timeout 900 python -m pytest tests -m gpu -x -q -k conv > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -5 gpurun_out/${T}_pytest.log
python scratch/conv_probe.py f16x3 > gpurun_out/${T}_conv_probe.json 2> gpurun_out/${T}_conv_probe.err
cat gpurun_out/${T}_conv_probe.json
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_conv_launches.csv python scratch/conv_probe.py f16x3 2 > gpurun_out/${T}_conv_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'conv2_refine' -s 1 -c 1 -o gpurun_out/${T}_refine_full python scratch/conv_probe.py f16x3 1 > gpurun_out/${T}_refine_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'fc_fused_kernel' -s 6 -c 1 -o gpurun_out/${T}_fused2p_full python bench.py --no-cpu-baseline --no-extra --steps 1 --warmup 3 > gpurun_out/${T}_fused_full.log 2>&1
tail -2 gpurun_out/${T}_fused_full.log
RBNN_FUSED_DEBUG=8 python bench.py --no-cpu-baseline --no-extra --steps 1 --warmup 3 --samples 334 2>&1 | grep "fused cta0" | tail -4 > gpurun_out/${T}_timers.log
cat gpurun_out/${T}_timers.log
