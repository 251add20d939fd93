#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/flr_tile_sweep.py > gpurun_out/flr_tile_sweep.txt 2>&1; echo "sweep rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_tc5_kernel -s 1 -c 5 -o gpurun_out/prof_wgrad_tc5 -f python tools/wgrad_prof.py > gpurun_out/ncu_wgrad.log 2>&1; echo "ncu rc=$?"
cat gpurun_out/flr_tile_sweep.txt
