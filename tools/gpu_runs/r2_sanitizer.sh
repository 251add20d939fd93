#!/bin/bash
# compute-sanitizer over the small-shape paths: memcheck of smoke() (tiny generator: fp32, tcgen05 and fast inference paths, one
# training step), then memcheck of a few operator tests with ragged shapes.  Summaries only (gpurun_out/ is size limited).
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 --error-exitcode 7 python __graft_entry__.py smoke > gpurun_out/san_smoke.log 2>&1; echo "memcheck smoke rc=$?"
grep -E "ERROR SUMMARY|Invalid|out of bounds|misaligned|smoke:" gpurun_out/san_smoke.log | head -20
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 --error-exitcode 7 python -m pytest tests/test_gpu_tc.py -x -q -k "ragged or shape or odd or pitched or hybrid" > gpurun_out/san_tc.log 2>&1; echo "memcheck tc tests rc=$?"
grep -E "ERROR SUMMARY|Invalid|out of bounds|misaligned|passed|failed" gpurun_out/san_tc.log | head -20
