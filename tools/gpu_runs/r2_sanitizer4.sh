#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 compute-sanitizer --tool racecheck --print-limit 10 --error-exitcode 7 python -m pytest tests/test_gpu_tc.py -x -q -k "bit_identical and (shape0 or shape2 or shape8)" > gpurun_out/race_tc2.log 2>&1; echo "racecheck rc=$?"
grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/race_tc2.log | head
tail -c 2000 gpurun_out/race_tc2.log > gpurun_out/race_tc2_tail.log; rm -f gpurun_out/race_tc2.log
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_generator.py -x -q 2>&1 | tail -2
