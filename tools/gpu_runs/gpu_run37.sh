#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/summary37.txt; : > $S
timeout 900 python -m pytest tests/test_gpu_flr_tcs.py -q -s 2>&1 | grep -E "fwd|^E   +Assert|passed|failed|FAILED" | cut -c1-250 | head -40 >> $S
timeout 900 python bench.py --workload train --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_train_n1.log 2> gpurun_out/bench_train_n1.err; echo "bench_train rc=$?" >> $S
cat $S; tail -3 gpurun_out/bench_train_n1.err; python - <<'P'
import json
d=json.loads(open('gpurun_out/bench_train_n1.log').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['config']['first_losses'])
for k,v in d['rooflines'].items():
    if v: print(k, round(v['ms_per_step'],2), round(v['achieved'],1), v['unit'], round(v['frac'],3))
P
