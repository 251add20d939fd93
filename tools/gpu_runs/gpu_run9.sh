#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/summary.txt; : > $S
timeout 600 python -m pytest tests/test_gpu_flr_tc.py -m gpu -q -x > gpurun_out/t_flr_tc.log 2>&1; echo "flr_tc tests rc=$?" >> $S
timeout 600 python tools/layer_bench.py --batch 16 --ops flrelu_tc --json gpurun_out/lb_tc.json > gpurun_out/lb_tc.log 2>&1; echo "lb rc=$?" >> $S
timeout 600 python tools/layer_bench.py --batch 16 --ops flrelu_tc,f16in,f16out > gpurun_out/lb_tc16.log 2>&1; echo "lb16 rc=$?" >> $S
timeout 900 ncu --set full --clock-control none --import-source on -k regex:flr_tc_kernel -s 10 -c 3 -o gpurun_out/prof_flr_tc -f \
    python tools/layer_bench.py --batch 16 --ops flrelu_tc > gpurun_out/ncu_flr_tc.log 2>&1; echo "ncu rc=$?" >> $S
cat $S; tail -5 gpurun_out/t_flr_tc.log; grep SUMMARY gpurun_out/lb_tc.log gpurun_out/lb_tc16.log
