#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:flr_tcs_kernel -s 2 -c 2 -o gpurun_out/prof_flr_tcs -f python tools/flr_prof.py tc > gpurun_out/ncu_flr_tcs.log 2>&1; echo "rc=$?"
tail -3 gpurun_out/ncu_flr_tcs.log
