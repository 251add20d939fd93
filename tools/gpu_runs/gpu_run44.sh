#!/bin/bash
# A/B: narrow last strip of the tensor-core filtered_lrelu against the committed build (same box)
mkdir -p gpurun_out
S=gpurun_out/summary44.txt; : > $S
timeout 600 python -m pytest tests/test_gpu_flr_tc.py tests/test_gpu_generator.py -m gpu -q -x 2>&1 | tail -5 >> $S
OPS=flrelu_tc,f16in,f16out,nobias
AFCM_B200_LIB=$PWD/afcm_b200/libafcm_b200_prev.so timeout 300 python tools/layer_bench.py --batch 64 --ops $OPS --json gpurun_out/flr_a.json > gpurun_out/flr_a.log 2>&1; echo "a rc=$?" >> $S
timeout 300 python tools/layer_bench.py --batch 64 --ops $OPS --json gpurun_out/flr_b.json > gpurun_out/flr_b.log 2>&1; echo "b rc=$?" >> $S
AFCM_B200_LIB=$PWD/afcm_b200/libafcm_b200_prev.so timeout 300 python tools/layer_bench.py --batch 64 --ops $OPS --json gpurun_out/flr_a2.json > gpurun_out/flr_a2.log 2>&1; echo "a2 rc=$?" >> $S
timeout 300 python tools/layer_bench.py --batch 64 --ops $OPS --json gpurun_out/flr_b2.json > gpurun_out/flr_b2.log 2>&1; echo "b2 rc=$?" >> $S
cat $S; for f in a b a2 b2; do grep SUMMARY gpurun_out/flr_$f.log | grep -o '"flrelu_tc_ms": [0-9.]*'; done
