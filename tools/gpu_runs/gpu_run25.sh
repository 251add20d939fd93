#!/bin/bash
# full GPU test tier + default bench (both arms) on one B200
mkdir -p gpurun_out
S=gpurun_out/summary25.txt; : > $S
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 >> $S
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.log 2> gpurun_out/bench_n1.err; echo "bench_n1 rc=$?" >> $S
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "bench_ref rc=$?" >> $S
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" >> $S 2>&1
cat $S; cut -c1-600 gpurun_out/bench_n1.log
