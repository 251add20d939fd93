#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:flr_tc_kernel -o gpurun_out/prof_flr_v3 -f \
    python tools/layer_bench.py --batch 16 --ops flrelu_tc,f16in,f16out,nobias --layers enc3,enc4,enc12,L10 --iters 1 --warmup 1 > gpurun_out/ncu_flr_v3.log 2>&1; echo "ncu rc=$?"
