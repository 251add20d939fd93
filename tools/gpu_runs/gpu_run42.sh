#!/bin/bash
# waves sweep of the persistent tensor-core filtered_lrelu
mkdir -p gpurun_out
OPS=flrelu_tc,f16in,f16out,nobias
for w in 0 4 8 16 32; do
  AFCM_FTC_WAVES=$w timeout 300 python tools/layer_bench.py --batch 64 --ops $OPS --json gpurun_out/flr_w$w.json > gpurun_out/flr_w$w.log 2>&1
  echo "w=$w $(grep SUMMARY gpurun_out/flr_w$w.log | grep -o '"flrelu_tc_ms": [0-9.]*')"
done
