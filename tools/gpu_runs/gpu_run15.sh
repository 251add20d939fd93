#!/bin/bash
# full GPU tier + bench (graph) + 2-GPU torchrun bench
mkdir -p gpurun_out
S=gpurun_out/summary.txt; : > $S
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/t_gpu.log 2>&1; echo "pytest_gpu rc=$?" >> $S
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> $S
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench1.log 2> gpurun_out/bench1.err; echo "bench1 rc=$?" >> $S
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --graph 0 > gpurun_out/bench1_eager.log 2> gpurun_out/bench1e.err; echo "bench1 eager rc=$?" >> $S
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --batch 1 > gpurun_out/bench_b1.log 2> gpurun_out/bench_b1.err; echo "bench b1 rc=$?" >> $S
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --batch 1 --graph 0 > gpurun_out/bench_b1e.log 2> gpurun_out/bench_b1e.err; echo "bench b1 eager rc=$?" >> $S
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench2.log 2> gpurun_out/bench2.err; echo "bench2 rc=$?" >> $S
fi
cat $S; tail -4 gpurun_out/t_gpu.log; tail -1 gpurun_out/smoke.log
for f in bench1 bench1_eager bench_b1 bench_b1e bench2; do echo "== $f"; cut -c1-330 gpurun_out/$f.log 2>/dev/null; done; tail -3 gpurun_out/bench2.err 2>/dev/null
