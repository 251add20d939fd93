timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_cm_networks.py tests/test_gpu_train.py -x -q 2>&1 | tail -3
timeout 300 python tools/upfirdn_bench.py 2>&1 | tail -6
