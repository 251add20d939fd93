mkdir -p gpurun_out
timeout 600 python tools/flr_t5_check.py --big > gpurun_out/t5_check.txt 2>&1; echo "rc=$?" >> gpurun_out/t5_check.txt
tail -60 gpurun_out/t5_check.txt
