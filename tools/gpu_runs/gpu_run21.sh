#!/bin/bash
# training-step path: gradient parity tests
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -x -q 2>&1 | tail -40 > gpurun_out/train_tests.log; echo "rc=$?" >> gpurun_out/train_tests.log
cat gpurun_out/train_tests.log
