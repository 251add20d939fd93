#!/bin/bash
# two GPUs: forward bench under torchrun with the final code (pipelined e2e, persistent flr_tc)
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err; echo "bench_n2 rc=$?"
tail -2 gpurun_out/bench_n2.err; cut -c1-300 gpurun_out/bench_n2.log
