# correctness + per-layer timing + timeline of the tcgen05 filtered_lrelu (development iteration)
mkdir -p gpurun_out
B=${1:-32}
timeout 600 python tools/flr_t5_check.py --big > gpurun_out/t5_check.txt 2>&1; echo "check rc=$?"; grep -E "FAIL|ALL OK|error" gpurun_out/t5_check.txt | head
timeout 600 python tools/layer_bench.py --batch $B --ops flrelu_tc,f16in,f16out,t5 --json gpurun_out/lb_t5.json > gpurun_out/lb_t5.log 2>&1; echo "t5 rc=$?"
python - <<'PY'
import json
b=json.load(open('gpurun_out/lb_t5.json'))
for rb in b['rows']:
    if 'flrelu_tc_ms' in rb:
        print('%-14s C=%3d %3d->%3d u%d d%d  t5 %.3f ms %.3f' % (rb['layer'],rb['cout'],rb['Hc'],rb['out'],rb['up'],rb['down'],rb['flrelu_tc_ms'],rb['flrelu_tc_frac']))
print('T5', {k:v for k,v in b['summary'].items() if 'flrelu_tc' in k})
PY
timeout 300 python tools/flr_t5_trace.py --size 278 --planes 2048 --steps 2 > gpurun_out/t5_trace_278.txt 2>&1; echo "rc=$?"
cat gpurun_out/t5_trace_278.txt
