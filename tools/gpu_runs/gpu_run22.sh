#!/bin/bash
# training-step path: gradient parity tests + first training-step bench line (N=1)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -x -q 2>&1 | tail -15 > gpurun_out/train_tests.log; echo "rc=$?" >> gpurun_out/train_tests.log
timeout 900 python bench.py --workload train --steps 3 --warmup 3 > gpurun_out/bench_train_n1.log 2> gpurun_out/bench_train_n1.err; echo "bench_train rc=$?" >> gpurun_out/train_tests.log
cat gpurun_out/train_tests.log; tail -5 gpurun_out/bench_train_n1.err; cut -c1-3000 gpurun_out/bench_train_n1.log
