#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/tc_dbg.log
for shape in "1 64 64 14 14 2" "2 128 96 36 36 2" "1 4 64 52 52 2"; do
  echo "=== shape $shape" >> gpurun_out/tc_dbg.log
  TC_DBG_MODE=0 timeout 120 python tools/tc_debug.py $shape >> gpurun_out/tc_dbg.log 2>&1
  echo "tcdbg [$shape] rc=$?" >> gpurun_out/summary.txt
done
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q -s > gpurun_out/t_tc.log 2>&1
echo "tc rc=$?" >> gpurun_out/summary.txt
timeout 600 python tools/layer_bench.py --batch 8 --ops conv_tc,conv_f32 --json gpurun_out/lb_conv.json > gpurun_out/lb_conv.log 2>&1
echo "lb_conv rc=$?" >> gpurun_out/summary.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; cat gpurun_out/tc_dbg.log; tail -25 gpurun_out/t_tc.log; tail -3 gpurun_out/smoke.log; tail -3 gpurun_out/lb_conv.log
