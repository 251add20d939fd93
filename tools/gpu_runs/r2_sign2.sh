mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_flr_tcs.py tests/test_gpu_train.py -q 2>&1 | tail -4
