#!/bin/bash
# ldmatrix / stmatrix producers of the direct convolution on pitched planes: parity, repeatability, per-layer times
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_generator.py -x -q 2>&1 | tail -3
timeout 600 python tools/layer_bench.py --batch 32 --ops conv_tc,conv_nchw,f16in,f16out --json gpurun_out/lb_pitched.json > gpurun_out/lb_pitched.log 2>&1
python - <<'PY'
import json
a=json.load(open('gpurun_out/lb_pitched.json')); rows=a['rows'] if isinstance(a,dict) else a
t=[0,0,0]
for x in rows:
    if 'conv_tc_ms' in x:
        print('%-14s %3d %3d %3d packed %.3f dense-direct %.3f pitched-direct %.3f' % (x['layer'], x['cin'], x['cout'], x['H'], x['conv_tc_ms'], x.get('conv_nchw_ms',0), x.get('conv_pitched_ms',0)))
        t[0]+=x['conv_tc_ms']; t[1]+=x.get('conv_nchw_ms',0); t[2]+=x.get('conv_pitched_ms',0)
print('totals', t)
PY
