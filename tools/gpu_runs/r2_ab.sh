#!/bin/bash
# A/B of two builds of the library on one box: forward bench, alternating, three rounds.
# Before the gpurun call: build the other revision (git stash / checkout, python -m afcm_b200.build) and copy its library to
# afcm_b200/libafcm_b200_prev.so (in-tree .so files travel to the box; AFCM_B200_LIB selects the library).
mkdir -p gpurun_out
for r in 1 2 3; do
  for v in prev new; do
    if [ $v = prev ]; then export AFCM_B200_LIB=$PWD/afcm_b200/libafcm_b200_prev.so; else unset AFCM_B200_LIB; fi
    timeout 600 python bench.py --steps 10 --warmup 3 --no-fp32-leg --no-cpu-baseline > gpurun_out/ab_$v$r.json 2> gpurun_out/ab_$v$r.err
    python - <<PY
import json
b=json.loads(open("gpurun_out/ab_$v$r.json").read().strip().splitlines()[-1])
print("$v$r", round(b["value"],1), round(b["ms_per_step"],2), "conv", round(b["rooflines"]["conv2d_tc"]["ms_per_step"],2), "flr", round(b["rooflines"]["filtered_lrelu"]["ms_per_step"],2), b["clocks"]["sm_mhz"])
PY
  done
done
