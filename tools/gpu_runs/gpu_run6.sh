#!/bin/bash
# bench line + ncu launch list + full captures (baseline before the filtered_lrelu rewrite)
mkdir -p gpurun_out
S=gpurun_out/summary.txt; : > $S
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> $S
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu_launches rc=$?" >> $S
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv2d_tc_kernel -s 40 -c 3 -o gpurun_out/prof_conv -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --batch 16 > gpurun_out/ncu_conv.log 2>&1; echo "ncu_conv rc=$?" >> $S
timeout 900 ncu --set full --clock-control none --import-source on -k regex:flr_fused -s 40 -c 4 -o gpurun_out/prof_flr -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --batch 16 > gpurun_out/ncu_flr.log 2>&1; echo "ncu_flr rc=$?" >> $S
cat $S; cat gpurun_out/bench.log; tail -3 gpurun_out/bench.err
