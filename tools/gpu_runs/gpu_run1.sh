#!/bin/bash
# First GPU session: parity tests, tcgen05 tests in their own process (a trap must not poison the rest),
# then per-layer timings.  Everything is logged under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python -m pytest tests/test_gpu_ops.py tests/test_gpu_generator.py -m gpu -q -x > gpurun_out/t_ops.log 2>&1
echo "ops rc=$?" >> gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q -s > gpurun_out/t_tc.log 2>&1
echo "tc rc=$?" >> gpurun_out/summary.txt
timeout 600 python tools/layer_bench.py --batch 8 --ops flrelu --json gpurun_out/lb_flrelu.json > gpurun_out/lb_flrelu.log 2>&1
echo "lb_flrelu rc=$?" >> gpurun_out/summary.txt
timeout 600 python tools/layer_bench.py --batch 8 --ops conv_tc,conv_f32 --json gpurun_out/lb_conv.json > gpurun_out/lb_conv.log 2>&1
echo "lb_conv rc=$?" >> gpurun_out/summary.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -5 gpurun_out/t_ops.log; tail -15 gpurun_out/t_tc.log
