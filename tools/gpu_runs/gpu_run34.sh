#!/bin/bash
# compute-sanitizer memcheck over the gradient kernels and the restructured filtered_lrelu
mkdir -p gpurun_out
S=gpurun_out/summary34.txt; : > $S
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest tests/test_gpu_train.py -q -x -k "modulated or conv_grads_vs_torch or fc_grads or tiny_generator_grads_fp32 or trainer" > gpurun_out/memcheck_train.log 2>&1; echo "memcheck train rc=$?" >> $S
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest tests/test_gpu_train.py -q -x -k "conv_grads_tc_vs_fp32 and bfloat16" > gpurun_out/memcheck_tc.log 2>&1; echo "memcheck tc rc=$?" >> $S
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest tests/test_gpu_ops.py -q -x -k "filtered_lrelu" > gpurun_out/memcheck_flr.log 2>&1; echo "memcheck flr rc=$?" >> $S
cat $S; tail -4 gpurun_out/memcheck_train.log; tail -4 gpurun_out/memcheck_tc.log; tail -4 gpurun_out/memcheck_flr.log; grep -h -A8 "Invalid\|out of bounds\|misaligned" gpurun_out/memcheck_*.log | head -40
