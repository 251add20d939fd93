#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/summary30.txt; : > $S
timeout 1500 python -m pytest tests/test_gpu_ops.py tests/test_gpu_train.py tests/test_gpu_generator.py -m gpu -q 2>&1 | tail -4 >> $S
timeout 900 python bench.py --workload train --steps 3 --warmup 3 > gpurun_out/bench_train_n1.log 2> gpurun_out/bench_train_n1.err; echo "bench_train rc=$?" >> $S
timeout 900 ncu --set full --clock-control none --import-source on -k regex:flr_fused_kernel -s 2 -c 2 -o gpurun_out/prof_flr_simt -f python tools/flr_prof.py > gpurun_out/ncu_flr_simt.log 2>&1; echo "ncu rc=$?" >> $S
cat $S; python - <<'P'
import json
d=json.loads(open('gpurun_out/bench_train_n1.log').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'])
for k,v in d['rooflines'].items():
    if v: print(k, round(v['ms_per_step'],2), round(v['achieved'],1), v['unit'], round(v['frac'],3))
P
