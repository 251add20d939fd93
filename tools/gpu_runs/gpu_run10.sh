#!/bin/bash
# re-entry session: confirm the GPU state of the tree, per-layer flr_tc table (fp32 and fp16 I/O), ncu full of flr_tc
# on one layer of each geometry
mkdir -p gpurun_out
S=gpurun_out/summary.txt; : > $S
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/t_gpu.log 2>&1; echo "pytest_gpu rc=$?" >> $S
timeout 600 python tools/layer_bench.py --batch 16 --ops flrelu_tc --json gpurun_out/lb_tc32.json > gpurun_out/lb_tc32.log 2>&1; echo "lb32 rc=$?" >> $S
timeout 600 python tools/layer_bench.py --batch 16 --ops flrelu_tc,f16in,f16out --json gpurun_out/lb_tc16.json > gpurun_out/lb_tc16.log 2>&1; echo "lb16 rc=$?" >> $S
timeout 900 ncu --set full --clock-control none --import-source on -k regex:flr_tc_kernel -o gpurun_out/prof_flr_tc32 -f \
    python tools/layer_bench.py --batch 16 --ops flrelu_tc --layers enc0,enc4,enc12,L10,L3 --iters 1 --warmup 1 > gpurun_out/ncu_flr_tc32.log 2>&1; echo "ncu32 rc=$?" >> $S
timeout 900 ncu --set full --clock-control none --import-source on -k regex:flr_tc_kernel -o gpurun_out/prof_flr_tc16 -f \
    python tools/layer_bench.py --batch 16 --ops flrelu_tc,f16in,f16out --layers enc0,enc4,enc12,L10,L3 --iters 1 --warmup 1 > gpurun_out/ncu_flr_tc16.log 2>&1; echo "ncu16 rc=$?" >> $S
cat $S; tail -5 gpurun_out/t_gpu.log; grep SUMMARY gpurun_out/lb_tc32.log gpurun_out/lb_tc16.log
