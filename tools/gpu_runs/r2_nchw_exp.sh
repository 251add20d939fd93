mkdir -p gpurun_out
for dbg in 0 256 448; do
AFCM_TC_DBG=$dbg timeout 600 python tools/layer_bench.py --batch 64 --ops conv_tc,conv_nchw,f16in,f16out --layers enc1,enc3,enc7,L11 --json gpurun_out/lb_nchw_exp.json > gpurun_out/lb_nchw_exp.log 2>&1; echo "dbg=$dbg lb rc=$?"
python - <<'PY'
import json
b=json.load(open('gpurun_out/lb_nchw_exp.json'))
for r in b['rows']:
    if 'conv_tc_ms' in r:
        print('%-12s %3d->%3d @%3d pack %.3f conv %.3f  direct %.3f  pitched %.3f' % (r['layer'],r['cin'],r['cout'],r['H'],r['pack_ms'],r['conv_tc_ms'],r.get('conv_nchw_ms',0), r.get('conv_pitched_ms',0)))
PY
done
