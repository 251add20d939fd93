#!/bin/bash
# one or two MMA-issuing warps in the tcgen05 convolution: timeline of CTA 0, parity tests, per-layer times
mkdir -p gpurun_out
timeout 200 python tools/conv_tc_trace.py --cin 64 --cout 64 --issuers 2 --tiles 2 > gpurun_out/trace_iss2.txt 2>&1
timeout 200 python tools/conv_tc_trace.py --cin 64 --cout 64 --issuers 2 --tiles 2 --direct 1 > gpurun_out/trace_iss2_direct.txt 2>&1
timeout 200 python tools/conv_tc_trace.py --cin 128 --cout 128 --issuers 2 --tiles 1 > gpurun_out/trace_128.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_generator.py -x -q 2>&1 | tail -3
for i in 1 2; do
AFCM_TC_ISSUERS=$i timeout 600 python tools/layer_bench.py --batch 32 --ops conv_tc,conv_nchw,f16in,f16out --json gpurun_out/lb_iss$i.json > gpurun_out/lb_iss$i.log 2>&1
done
tail -42 gpurun_out/trace_iss2.txt
