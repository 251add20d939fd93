#!/bin/bash
# final-state profiles on one GPU: launch list with DRAM bytes per launch, small full captures of the two top kernels
mkdir -p gpurun_out
S=gpurun_out/summary.txt; : > $S
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.log 2> gpurun_out/bench_n1.err; echo "bench_n1 rc=$?" >> $S
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "bench_ref rc=$?" >> $S
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --graph 0 > gpurun_out/ncu_bench.log 2>&1; echo "ncu_launches rc=$?" >> $S
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv2d_tc_kernel -s 36 -c 4 -o gpurun_out/prof_conv_b64 -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --graph 0 > gpurun_out/ncu_conv.log 2>&1; echo "ncu_conv rc=$?" >> $S
timeout 900 ncu --set full --clock-control none --import-source on -k regex:flr_tc_kernel -s 30 -c 5 -o gpurun_out/prof_flr_b64 -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --graph 0 > gpurun_out/ncu_flr.log 2>&1; echo "ncu_flr rc=$?" >> $S
cat $S; cut -c1-300 gpurun_out/bench_n1.log; ls -la gpurun_out
