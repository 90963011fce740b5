T=${1:-s3m}
timeout 1700 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "(conv and not cfg4) or fc2 or two_pass or (tcgen05 and not conv) or golden_svi or halfmoons or half_moons" > gpurun_out/${T}_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/${T}_memcheck.log
grep -E "ERROR SUMMARY|passed|failed|memcheck rc|Invalid|out of bounds" gpurun_out/${T}_memcheck.log | tail -12
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "test_tcgen05_conv_engine_vs_oracle or (test_tcgen05_engine_vs_oracle and fc2)" > gpurun_out/${T}_racecheck.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/${T}_racecheck.log
grep -E "RACECHECK SUMMARY|passed|failed|racecheck rc|hazard" gpurun_out/${T}_racecheck.log | tail -8
