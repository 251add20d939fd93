#!/bin/bash
# pipelined host-to-host runner: test + bench e2e
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_generator.py -m gpu -q -x 2>&1 | tail -5
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.log 2> gpurun_out/bench_n1.err; echo "bench_n1 rc=$?"; tail -3 gpurun_out/bench_n1.err
python -c "
import json
d=json.load(open('gpurun_out/bench_n1.log')); print(d['value'], d['ms_per_step'], d['e2e'])"
