#!/bin/bash
# final-state measurements: full GPU tier, smoke, bench N=1 (with CPU baseline), reference arm, launch list, ncu full of the
# conv and flr kernels in the bench configuration, volume bench at 1 and 2 GPUs, bench N=2
mkdir -p gpurun_out
S=gpurun_out/summary.txt; : > $S
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/t_gpu.log 2>&1; echo "pytest_gpu rc=$?" >> $S
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> $S
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.log 2> gpurun_out/bench_n1.err; echo "bench_n1 rc=$?" >> $S
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "bench_ref rc=$?" >> $S
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --graph 0 > gpurun_out/ncu_bench.log 2>&1; echo "ncu_launches rc=$?" >> $S
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv2d_tc_kernel -s 29 -c 29 -o gpurun_out/prof_conv_b64 -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --graph 0 > gpurun_out/ncu_conv.log 2>&1; echo "ncu_conv rc=$?" >> $S
timeout 900 ncu --set full --clock-control none --import-source on -k regex:flr_tc_kernel -s 28 -c 28 -o gpurun_out/prof_flr_b64 -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --graph 0 > gpurun_out/ncu_flr.log 2>&1; echo "ncu_flr rc=$?" >> $S
timeout 600 python tools/volume_bench.py --slices 160 --thickness 5 --batch 32 > gpurun_out/vol_n1.log 2> gpurun_out/vol_n1.err; echo "vol_n1 rc=$?" >> $S
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/volume_bench.py --slices 160 --thickness 5 --batch 32 > gpurun_out/vol_n2.log 2> gpurun_out/vol_n2.err; echo "vol_n2 rc=$?" >> $S
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err; echo "bench_n2 rc=$?" >> $S
fi
cat $S; tail -3 gpurun_out/t_gpu.log; tail -1 gpurun_out/smoke.log
for f in bench_n1 bench_ref vol_n1 vol_n2 bench_n2; do echo "== $f"; cut -c1-400 gpurun_out/$f.log 2>/dev/null; done
