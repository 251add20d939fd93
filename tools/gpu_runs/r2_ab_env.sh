#!/bin/bash
# A/B of tuning switches on one box: forward bench, alternating, two rounds per setting
mkdir -p gpurun_out
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-fp32-leg --no-cpu-baseline > gpurun_out/abe_$name.json 2> gpurun_out/abe_$name.err
  python - <<PY
import json
b=json.loads(open("gpurun_out/abe_$name.json").read().strip().splitlines()[-1])
print("$name", round(b["value"],1), round(b["ms_per_step"],2), "conv", round(b["rooflines"]["conv2d_tc"]["ms_per_step"],2), "flr", round(b["rooflines"]["filtered_lrelu"]["ms_per_step"],2), "pack", round(b["rooflines"]["conv_tc_pack"]["ms_per_step"],2), b["clocks"]["sm_mhz"])
PY
}
for r in 1 2; do
  run base$r X=1
  run iss2_$r AFCM_TC_ISSUERS=2
  run hybrid$r AFCM_HYBRID_PACK=1
  run both$r AFCM_HYBRID_PACK=1 AFCM_TC_ISSUERS=2
done
