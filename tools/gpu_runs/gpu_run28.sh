#!/bin/bash
# tcgen05 weight gradient: parity first (own process: a trap poisons the context), then per-layer timing, then the train bench
mkdir -p gpurun_out
S=gpurun_out/summary28.txt; : > $S
timeout 600 python -m pytest tests/test_gpu_train.py -x -q -k "conv_grads_tc_vs_fp32" 2>&1 | tail -25 >> $S
timeout 600 python -m pytest tests/test_gpu_train.py -q 2>&1 | tail -6 >> $S
timeout 600 python tools/wgrad_bench.py 2>&1 | tail -31 >> $S
timeout 900 python bench.py --workload train --steps 3 --warmup 3 > gpurun_out/bench_train_n1.log 2> gpurun_out/bench_train_n1.err; echo "bench_train rc=$?" >> $S
cat $S; python - <<'P'
import json
d=json.loads(open('gpurun_out/bench_train_n1.log').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'])
for k,v in d['rooflines'].items():
    if v: print(k, round(v['ms_per_step'],2), round(v['achieved'],1), v['unit'], round(v['frac'],3))
P
