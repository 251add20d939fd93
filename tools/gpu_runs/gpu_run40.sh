#!/bin/bash
# re-validation after the container rebuild (one GPU): full GPU test tier, smoke, both bench arms, training bench, volume
# throughput (configs 2-3), and ncu --set full captures (with source) of flr_tc (forward) and flr_tcs (training) for the
# warp-state breakdown that plans round 2.
mkdir -p gpurun_out
S=gpurun_out/summary40.txt; : > $S
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -5 >> $S
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" >> $S 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.log 2> gpurun_out/bench_n1.err; echo "bench_n1 rc=$?" >> $S
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flr_tc_kernel -s 1 -c 7 -o gpurun_out/prof_flr_tc_r1 -f python tools/flr_tc_prof.py > gpurun_out/ncu_flr_tc.log 2>&1; echo "ncu_flr_tc rc=$?" >> $S
timeout 600 python tools/volume_bench.py > gpurun_out/volume_bench.log 2> gpurun_out/volume_bench.err; echo "volume rc=$?" >> $S
timeout 600 python bench.py --workload train --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_train_n1.log 2> gpurun_out/bench_train_n1.err; echo "bench_train rc=$?" >> $S
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flr_tcs_kernel -s 2 -c 2 -o gpurun_out/prof_flr_tcs_r1 -f python tools/flr_prof.py tc > gpurun_out/ncu_flr_tcs.log 2>&1; echo "ncu_tcs rc=$?" >> $S
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "bench_ref rc=$?" >> $S
cat $S; cut -c1-300 gpurun_out/bench_n1.log; cut -c1-200 gpurun_out/bench_train_n1.log; tail -3 gpurun_out/volume_bench.log
