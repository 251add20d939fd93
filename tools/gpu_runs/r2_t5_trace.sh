mkdir -p gpurun_out
timeout 300 python tools/flr_t5_trace.py --size 278 --planes 2048 > gpurun_out/t5_trace_278.txt 2>&1; echo "rc=$?"
timeout 300 python tools/flr_t5_trace.py --size 38 --planes 8192 --steps 3 > gpurun_out/t5_trace_38.txt 2>&1; echo "rc=$?"
cat gpurun_out/t5_trace_278.txt; cat gpurun_out/t5_trace_38.txt
