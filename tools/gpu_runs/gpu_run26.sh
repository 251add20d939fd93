#!/bin/bash
# full GPU test tier (no -x) + training bench
mkdir -p gpurun_out
S=gpurun_out/summary26.txt; : > $S
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12 >> $S
timeout 900 python bench.py --workload train --steps 3 --warmup 3 > gpurun_out/bench_train_n1.log 2> gpurun_out/bench_train_n1.err; echo "bench_train rc=$?" >> $S
cat $S; python - <<'P'
import json
d=json.loads(open('gpurun_out/bench_train_n1.log').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'])
for k,v in d['rooflines'].items():
    if v: print(k, round(v['ms_per_step'],2), round(v['achieved'],1), v['unit'], round(v['frac'],3))
P
