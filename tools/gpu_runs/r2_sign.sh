mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_flr_tcs.py -x -q -s 2>&1 | grep -v "^$" | tail -40
