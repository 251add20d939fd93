#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv2d_tc_kernel -o gpurun_out/prof_conv_small -f \
    python tools/layer_bench.py --batch 16 --ops conv_tc,f16in,f16out --layers enc1,enc3,enc8 --iters 1 --warmup 1 > gpurun_out/ncu_conv_small.log 2>&1; echo "ncu rc=$?"
