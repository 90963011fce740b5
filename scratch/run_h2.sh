T=${1:-s4e}
timeout 2400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "not headline and not cfg4 and not two_pass" > gpurun_out/${T}_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/${T}_memcheck.log
grep -E "ERROR SUMMARY|passed|failed|memcheck rc" gpurun_out/${T}_memcheck.log | tail -4
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "two_pass and not 1000" > gpurun_out/${T}_memcheck2.log 2>&1
echo "memcheck2 rc=$?" >> gpurun_out/${T}_memcheck2.log
grep -E "ERROR SUMMARY|passed|failed|memcheck2 rc" gpurun_out/${T}_memcheck2.log | tail -4
