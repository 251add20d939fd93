#!/bin/bash
# flr_tc v3 (FAST variant, cheaper edges): parity, per-layer table at 5 and 4 CTAs/SM, bench
mkdir -p gpurun_out
S=gpurun_out/summary.txt; : > $S
timeout 900 python -m pytest tests/test_gpu_flr_tc.py tests/test_gpu_generator.py -m gpu -q -x -s > gpurun_out/t_gpu.log 2>&1; echo "pytest_gpu rc=$?" >> $S
timeout 600 python tools/layer_bench.py --batch 16 --ops flrelu_tc,f16in,f16out,nobias --json gpurun_out/lb_fast5.json > gpurun_out/lb_fast5.log 2>&1; echo "lb5 rc=$?" >> $S
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench5.log 2> gpurun_out/bench.err; echo "bench5 rc=$?" >> $S
AFCM_NVCC_EXTRA=-DAFCM_FTC_MINB22=4 python -m afcm_b200.build > gpurun_out/build4.log 2>&1; echo "build4 rc=$?" >> $S
timeout 600 env AFCM_NVCC_EXTRA=-DAFCM_FTC_MINB22=4 python tools/layer_bench.py --batch 16 --ops flrelu_tc,f16in,f16out,nobias --json gpurun_out/lb_fast4.json > gpurun_out/lb_fast4.log 2>&1; echo "lb4 rc=$?" >> $S
cat $S; tail -8 gpurun_out/t_gpu.log; cat gpurun_out/bench5.log | cut -c1-400; grep SUMMARY gpurun_out/lb_fast5.log gpurun_out/lb_fast4.log
