mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/t5_probe.txt 2>&1
timeout 300 tools/microbench/t5_probe.bin >> gpurun_out/t5_probe.txt 2>&1; echo "rc=$?" >> gpurun_out/t5_probe.txt
cat gpurun_out/t5_probe.txt
