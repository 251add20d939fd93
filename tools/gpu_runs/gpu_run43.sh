#!/bin/bash
# validation of the persistent (16 waves) + trimmed tensor-core filtered_lrelu: full GPU tier, bench
mkdir -p gpurun_out
S=gpurun_out/summary43.txt; : > $S
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 >> $S
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.log 2> gpurun_out/bench_n1.err; echo "bench_n1 rc=$?" >> $S
timeout 300 python tools/volume_bench.py > gpurun_out/volume_bench.log 2> gpurun_out/volume_bench.err; echo "volume rc=$?" >> $S
cat $S; cut -c1-250 gpurun_out/bench_n1.log; cut -c1-200 gpurun_out/volume_bench.log
