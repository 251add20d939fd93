#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_ops.py tests/test_gpu_generator.py -m gpu -q > gpurun_out/t_ops.log 2>&1
echo "ops rc=$?" > gpurun_out/summary.txt
for mode in 8 9 10 12 0; do
  echo "=== TC_DBG_MODE=$mode" >> gpurun_out/tc_dbg.log
  TC_DBG_MODE=$mode timeout 120 python tools/tc_debug.py 1 64 64 14 14 2 >> gpurun_out/tc_dbg.log 2>&1
  echo "tcdbg mode $mode rc=$?" >> gpurun_out/summary.txt
done
cat gpurun_out/summary.txt; tail -8 gpurun_out/t_ops.log; cat gpurun_out/tc_dbg.log
