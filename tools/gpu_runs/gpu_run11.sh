#!/bin/bash
# fast path (fp16 storage, flr_tc wired in): tests, smoke, bench line, layer bench of pack/conv in fp16 mode, launch list
mkdir -p gpurun_out
S=gpurun_out/summary.txt; : > $S
timeout 900 python -m pytest tests -m gpu -q -x -s > gpurun_out/t_gpu.log 2>&1; echo "pytest_gpu rc=$?" >> $S
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> $S
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> $S
timeout 600 python tools/layer_bench.py --batch 16 --ops conv_tc,f16in,f16out --json gpurun_out/lb_conv16.json > gpurun_out/lb_conv16.log 2>&1; echo "lb_conv16 rc=$?" >> $S
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu_launches rc=$?" >> $S
cat $S; tail -15 gpurun_out/t_gpu.log; tail -2 gpurun_out/smoke.log; cat gpurun_out/bench.log; tail -3 gpurun_out/bench.err; grep SUMMARY gpurun_out/lb_conv16.log
