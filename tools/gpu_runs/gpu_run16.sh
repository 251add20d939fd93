#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/summary.txt; : > $S
timeout 900 python -m pytest tests/test_gpu_flr_tc.py tests/test_gpu_generator.py -m gpu -q -x > gpurun_out/t_gpu.log 2>&1; echo "pytest_gpu rc=$?" >> $S
timeout 600 python tools/layer_bench.py --batch 16 --ops flrelu_tc,f16in,f16out,nobias --json gpurun_out/lb_fast.json > gpurun_out/lb_fast.log 2>&1; echo "lb rc=$?" >> $S
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> $S
cat $S; tail -4 gpurun_out/t_gpu.log; cut -c1-200 gpurun_out/bench.log; grep SUMMARY gpurun_out/lb_fast.log | cut -c1-400
