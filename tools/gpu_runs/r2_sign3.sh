mkdir -p gpurun_out
timeout 900 python tools/flr_train_bench.py 2>&1 | tail -20
