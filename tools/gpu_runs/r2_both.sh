bash tools/gpu_runs/r2_nchw.sh
bash tools/gpu_runs/r2_validate.sh
