#!/bin/bash
mkdir -p gpurun_out
timeout 2400 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest tests/test_gpu_train.py -q -x -k "conv_grads_tc_vs_fp32 or grads_tc" > gpurun_out/memcheck_tc.log 2>&1; echo "memcheck tc rc=$?"
tail -6 gpurun_out/memcheck_tc.log; grep -h -A10 "Invalid\|out of bounds\|misaligned" gpurun_out/memcheck_tc.log | head -40
