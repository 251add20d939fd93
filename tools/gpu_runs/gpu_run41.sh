#!/bin/bash
# A/B of the tensor-core filtered_lrelu scheduling (persistent warps + pipeline fill/drain trimming) against the previous build
mkdir -p gpurun_out
S=gpurun_out/summary41.txt; : > $S
timeout 600 python -m pytest tests/test_gpu_flr_tc.py tests/test_gpu_generator.py -m gpu -q -x 2>&1 | tail -5 >> $S
OPS=flrelu_tc,f16in,f16out,nobias
AFCM_B200_LIB=$PWD/afcm_b200/libafcm_b200_prev.so timeout 300 python tools/layer_bench.py --batch 64 --ops $OPS --json gpurun_out/flr_prev.json > gpurun_out/flr_prev.log 2>&1; echo "prev rc=$?" >> $S
timeout 300 python tools/layer_bench.py --batch 64 --ops $OPS --json gpurun_out/flr_new.json > gpurun_out/flr_new.log 2>&1; echo "new rc=$?" >> $S
AFCM_FTC_WAVES=0 timeout 300 python tools/layer_bench.py --batch 64 --ops $OPS --json gpurun_out/flr_new_w0.json > gpurun_out/flr_new_w0.log 2>&1; echo "new_w0 rc=$?" >> $S
AFCM_FTC_WAVES=2 timeout 300 python tools/layer_bench.py --batch 64 --ops $OPS --json gpurun_out/flr_new_w2.json > gpurun_out/flr_new_w2.log 2>&1; echo "new_w2 rc=$?" >> $S
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.log 2> gpurun_out/bench_n1.err; echo "bench_n1 rc=$?" >> $S
cat $S; for f in prev new new_w0 new_w2; do grep SUMMARY gpurun_out/flr_$f.log | cut -c1-400; done; cut -c1-200 gpurun_out/bench_n1.log
