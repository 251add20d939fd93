#!/bin/bash
# ncu --set full captures of the FINAL tensor-core filtered_lrelu kernels (forward flr_tc, training flr_tcs)
mkdir -p gpurun_out
timeout 100 ncu --set full --clock-control none --import-source on -k regex:flr_tc_kernel -s 1 -c 7 -o gpurun_out/prof_flr_tc_final -f python tools/flr_tc_prof.py > gpurun_out/ncu_flr_tc.log 2>&1; echo "ncu_flr_tc rc=$?"
timeout 60 ncu --set full --clock-control none --import-source on -k regex:flr_tcs_kernel -s 2 -c 2 -o gpurun_out/prof_flr_tcs_final -f python tools/flr_prof.py tc > gpurun_out/ncu_flr_tcs.log 2>&1; echo "ncu_tcs rc=$?"
