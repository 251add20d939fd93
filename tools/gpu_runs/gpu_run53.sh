#!/bin/bash
# final state of round 1 after the pipelined runners and the load-pass changes: full GPU tier, smoke, forward + training bench
mkdir -p gpurun_out
S=gpurun_out/summary53.txt; : > $S
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 >> $S
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" >> $S 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.log 2> gpurun_out/bench_n1.err; echo "bench_n1 rc=$?" >> $S
timeout 600 python bench.py --workload train --steps 3 --warmup 3 > gpurun_out/bench_train_n1.log 2> gpurun_out/bench_train_n1.err; echo "bench_train rc=$?" >> $S
cat $S; cut -c1-250 gpurun_out/bench_n1.log; cut -c1-200 gpurun_out/bench_train_n1.log
