#!/bin/bash
# flr SIMT restructure: op tests + training bench
mkdir -p gpurun_out
S=gpurun_out/summary27.txt; : > $S
timeout 1500 python -m pytest tests/test_gpu_ops.py tests/test_gpu_train.py tests/test_gpu_generator.py -m gpu -q 2>&1 | tail -6 >> $S
timeout 900 python bench.py --workload train --steps 3 --warmup 3 > gpurun_out/bench_train_n1.log 2> gpurun_out/bench_train_n1.err; echo "bench_train rc=$?" >> $S
cat $S; python - <<'P'
import json
d=json.loads(open('gpurun_out/bench_train_n1.log').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'])
for k,v in d['rooflines'].items():
    if v: print(k, round(v['ms_per_step'],2), round(v['achieved'],1), v['unit'], round(v['frac'],3))
P
