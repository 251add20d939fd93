#!/bin/bash
# pipelined volume predictor: test + volume bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_generator.py -m gpu -q -x 2>&1 | tail -3
timeout 300 python tools/volume_bench.py > gpurun_out/volume_bench.log 2> gpurun_out/volume_bench.err; echo "volume rc=$?"; tail -3 gpurun_out/volume_bench.err; cut -c1-220 gpurun_out/volume_bench.log
timeout 300 python tools/volume_bench.py --batch 64 --slices 256 | cut -c1-220
