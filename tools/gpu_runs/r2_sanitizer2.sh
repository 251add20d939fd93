#!/bin/bash
# compute-sanitizer memcheck over the whole GPU test tier, racecheck (shared-memory hazards) over the operator tests
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1700 compute-sanitizer --tool memcheck --print-limit 30 --error-exitcode 7 python -m pytest tests -m gpu -x -q > gpurun_out/san_all.log 2>&1; echo "memcheck all rc=$?"
grep -E "ERROR SUMMARY|Invalid|out of bounds|misaligned|passed|failed" gpurun_out/san_all.log | head -20
timeout 1200 compute-sanitizer --tool racecheck --print-limit 30 --error-exitcode 7 python -m pytest tests/test_gpu_flr_tc.py tests/test_gpu_flr_tcs.py tests/test_gpu_ops.py -x -q > gpurun_out/race_ops.log 2>&1; echo "racecheck ops rc=$?"
grep -E "RACECHECK SUMMARY|hazard|passed|failed" gpurun_out/race_ops.log | sort | uniq -c | sort -rn | head -20
tail -c 3000 gpurun_out/san_all.log > gpurun_out/san_all_tail.log; tail -c 6000 gpurun_out/race_ops.log > gpurun_out/race_ops_tail.log
rm -f gpurun_out/san_all.log gpurun_out/race_ops.log
