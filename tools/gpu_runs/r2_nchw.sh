mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -x -q -k "direct_nchw" 2>&1 | tail -15
timeout 600 python tools/layer_bench.py --batch 64 --ops conv_tc,conv_nchw,f16in,f16out --json gpurun_out/lb_nchw.json > gpurun_out/lb_nchw.log 2>&1; echo "lb rc=$?"
python - <<'PY'
import json
b=json.load(open('gpurun_out/lb_nchw.json'))
for r in b['rows']:
    if 'conv_tc_ms' in r:
        print('%-12s %3d->%3d @%3d pack %.3f conv %.3f  direct %.3f  pitched %.3f (%+.3f vs pack+conv)' % (r['layer'],r['cin'],r['cout'],r['H'],r['pack_ms'],r['conv_tc_ms'],r.get('conv_nchw_ms',0), r.get('conv_pitched_ms',0), r.get('conv_pitched_ms',0)-r['pack_ms']-r['conv_tc_ms']))
print({k:v for k,v in b['summary'].items() if 'conv' in k or 'pack' in k})
PY
