mkdir -p gpurun_out
timeout 1800 python tools/ref_cuda_bench.py --batch 16 --json gpurun_out/r02_ref_cuda_bench.json > gpurun_out/r02_ref_cuda_bench.log 2>&1; echo "ref bench rc=$?"
tail -c 800 gpurun_out/r02_ref_cuda_bench.log | grep -v "^{" 
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r02_ref_cuda_bench.json'))
    print('plugin built', d['filtered_lrelu_plugin_built'], 'build s', round(d['plugin_build_s'],1), d.get('plugin_error'))
    print('flr', {k:v for k,v in d['filtered_lrelu'].items() if k!='rows'})
    for r in d['filtered_lrelu']['rows']: print('   %-12s C=%3d %3d->%3d u%d d%d %.3f ms %.0f GB/s' % (r['layer'],r['C'],r['Hc'],r['out'],r['up'],r['down'],r['ms'],r['gbs']))
    print('gen', d['generator']); print('gen strict', d['generator_fp32_strict'])
except Exception as e: print('no json', e)
PY
