#!/bin/bash
# A/B of flr_tc builds (CTAs per SM of the up 2 / down 2 kernel) on one box: per-layer bench, fp16 planes, batch 64
mkdir -p gpurun_out
for r in 1 2; do
for v in base m5 m6; do
  if [ $v = base ]; then unset AFCM_B200_LIB; else export AFCM_B200_LIB=$PWD/afcm_b200/libafcm_b200_$v.so; fi
  timeout 600 python tools/layer_bench.py --batch 64 --ops flrelu_tc,f16in,f16out,nobias --json gpurun_out/abf_$v$r.json > gpurun_out/abf_$v$r.log 2>&1
  python - <<PY
import json
a=json.load(open("gpurun_out/abf_$v$r.json")); rows=a['rows'] if isinstance(a,dict) else a
u22=sum(x['flrelu_tc_ms'] for x in rows if 'flrelu_tc_ms' in x and x['up']==2 and x['down']==2)
tot=sum(x['flrelu_tc_ms'] for x in rows if 'flrelu_tc_ms' in x)
big=[x['flrelu_tc_ms'] for x in rows if x['layer'] in ('enc1','enc5','enc8','enc12')]
print("$v$r", 'u2d2 layers %.3f ms, all %.3f ms' % (u22, tot), ['%.3f' % b for b in big])
PY
done
done
