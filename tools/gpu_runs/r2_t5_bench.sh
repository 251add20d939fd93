mkdir -p gpurun_out
B=${1:-32}
timeout 600 python tools/layer_bench.py --batch $B --ops flrelu_tc,f16in,f16out,nobias --json gpurun_out/lb_tc.json > gpurun_out/lb_tc.log 2>&1; echo "tc rc=$?"
timeout 600 python tools/layer_bench.py --batch $B --ops flrelu_tc,f16in,f16out,t5 --json gpurun_out/lb_t5.json > gpurun_out/lb_t5.log 2>&1; echo "t5 rc=$?"
python - <<'PY'
import json
a=json.load(open('gpurun_out/lb_tc.json')); b=json.load(open('gpurun_out/lb_t5.json'))
for ra,rb in zip(a['rows'],b['rows']):
    if 'flrelu_tc_ms' in ra:
        print('%-14s C=%3d %3d->%3d u%d d%d  tc %.3f ms %.2f | t5 %.3f ms %.2f' % (ra['layer'],ra['cout'],ra['Hc'],ra['out'],ra['up'],ra['down'],ra['flrelu_tc_ms'],ra['flrelu_tc_frac'],rb['flrelu_tc_ms'],rb['flrelu_tc_frac']))
print('TC', {k:v for k,v in a['summary'].items() if 'flrelu_tc' in k})
print('T5', {k:v for k,v in b['summary'].items() if 'flrelu_tc' in k})
PY
