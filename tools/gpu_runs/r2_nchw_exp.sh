mkdir -p gpurun_out
for rr in -1 0; do
AFCM_TC_ROWREUSE=$rr timeout 600 python tools/layer_bench.py --batch 64 --ops conv_tc,f16in,f16out --layers enc0,enc1,enc2,enc3,enc5,L10,L11,L12,L13 --json gpurun_out/lb_nchw_exp.json > gpurun_out/lb_nchw_exp.log 2>&1; echo "rowreuse=$rr lb rc=$?"
python - <<'PY'
import json
b=json.load(open('gpurun_out/lb_nchw_exp.json'))
print(' '.join('%s %.3f' % (r['layer'].split('_')[0], r['conv_tc_ms']) for r in b['rows'] if 'conv_tc_ms' in r))
PY
done
