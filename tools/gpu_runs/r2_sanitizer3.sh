#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 compute-sanitizer --tool racecheck --print-limit 30 --error-exitcode 7 python -m pytest tests/test_gpu_train.py tests/test_gpu_tc.py tests/test_gpu_cm_networks.py -x -q > gpurun_out/race_tc.log 2>&1; echo "racecheck tc/train rc=$?"
grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/race_tc.log | head
grep -E "hazard" gpurun_out/race_tc.log | cut -c1-160 | sort | uniq -c | sort -rn | head -12
tail -c 4000 gpurun_out/race_tc.log > gpurun_out/race_tc_tail.log; rm -f gpurun_out/race_tc.log
