#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/summary.txt; : > $S
timeout 900 python -m pytest tests -m gpu -q -x -s > gpurun_out/t_gpu.log 2>&1; echo "pytest_gpu rc=$?" >> $S
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> $S
cat $S; tail -8 gpurun_out/t_gpu.log; cat gpurun_out/bench.log; tail -3 gpurun_out/bench.err
