#!/bin/bash
# final state: full GPU test tier, both bench arms, training bench, launch list of one training step
mkdir -p gpurun_out
S=gpurun_out/summary32.txt; : > $S
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5 >> $S
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.log 2> gpurun_out/bench_n1.err; echo "bench_n1 rc=$?" >> $S
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "bench_ref rc=$?" >> $S
timeout 900 python bench.py --workload train --steps 3 --warmup 3 > gpurun_out/bench_train_n1.log 2> gpurun_out/bench_train_n1.err; echo "bench_train rc=$?" >> $S
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 4200 -c 1500 --csv --log-file gpurun_out/launches_train.csv \
    python bench.py --workload train --steps 1 --warmup 3 > gpurun_out/ncu_train.log 2>&1; echo "ncu_train rc=$?" >> $S
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" >> $S 2>&1
cat $S; cut -c1-300 gpurun_out/bench_n1.log; cut -c1-300 gpurun_out/bench_train_n1.log
