#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_ops.py tests/test_gpu_generator.py -m gpu -q -x > gpurun_out/t_ops.log 2>&1
echo "ops rc=$?" > gpurun_out/summary.txt
timeout 120 python tools/tc_debug.py 1 64 64 14 14 2 > gpurun_out/tc_dbg1.log 2>&1
echo "tcdbg1 rc=$?" >> gpurun_out/summary.txt
timeout 120 python tools/tc_debug.py 1 64 64 30 30 2 > gpurun_out/tc_dbg2.log 2>&1
echo "tcdbg2 rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -4 gpurun_out/t_ops.log; cat gpurun_out/tc_dbg1.log; cat gpurun_out/tc_dbg2.log
