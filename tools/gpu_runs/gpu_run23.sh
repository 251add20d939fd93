#!/bin/bash
# training-step path: tests, bench line, ncu launch list of one training step
mkdir -p gpurun_out
S=gpurun_out/summary23.txt; : > $S
timeout 900 python -m pytest tests/test_gpu_train.py -x -q 2>&1 | tail -5 >> $S
timeout 900 python bench.py --workload train --steps 3 --warmup 3 > gpurun_out/bench_train_n1.log 2> gpurun_out/bench_train_n1.err; echo "bench_train rc=$?" >> $S
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_train.csv \
    python bench.py --workload train --steps 1 --warmup 3 > gpurun_out/ncu_train.log 2>&1; echo "ncu_train rc=$?" >> $S
cat $S; python - <<'P'
import json
d=json.loads(open('gpurun_out/bench_train_n1.log').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'])
for k,v in d['rooflines'].items():
    if v: print(k, round(v['ms_per_step'],2), round(v['achieved'],1), v['unit'], round(v['frac'],3))
P
