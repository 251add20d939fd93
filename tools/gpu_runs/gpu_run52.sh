#!/bin/bash
# flr_tcs: batched tile loads (pass 0): parity tests + training bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_flr_tcs.py tests/test_gpu_train.py -m gpu -q -x 2>&1 | tail -3
timeout 600 python bench.py --workload train --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_train_n1.log 2> gpurun_out/bench_train_n1.err; echo "bench_train rc=$?"; tail -2 gpurun_out/bench_train_n1.err
python -c "
import json
d=json.load(open('gpurun_out/bench_train_n1.log')); print(d['value'], d['ms_per_step'], d['clocks'])
for k,v in d['rooflines'].items():
    if v: print(k, round(v['ms_per_step'],2), round(v['achieved'],1), v['unit'], round(v['frac'],3))"
