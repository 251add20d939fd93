#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/summary.txt; : > $S
timeout 600 python -m pytest tests/test_gpu_flr_tc.py -m gpu -q -x > gpurun_out/t_flr_tc.log 2>&1; echo "flr_tc tests rc=$?" >> $S
timeout 600 python tools/layer_bench.py --batch 16 --ops flrelu_tc --json gpurun_out/lb_tc.json > gpurun_out/lb_tc.log 2>&1; echo "layer_bench rc=$?" >> $S
cat $S; tail -30 gpurun_out/t_flr_tc.log; grep -v '^{' gpurun_out/lb_tc.log | tail -5
