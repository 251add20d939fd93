#!/bin/bash
# 2 GPUs: NCCL training-step check (pytest), training bench at N=2, inference bench at N=2
mkdir -p gpurun_out
S=gpurun_out/summary33.txt; : > $S
timeout 600 python -m pytest tests/test_gpu_train.py -q -k "two_gpu" 2>&1 | tail -5 >> $S
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --workload train --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_train_n2.log 2> gpurun_out/bench_train_n2.err; echo "bench_train_n2 rc=$?" >> $S
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err; echo "bench_n2 rc=$?" >> $S
cat $S; cut -c1-330 gpurun_out/bench_train_n2.log; cut -c1-330 gpurun_out/bench_n2.log
