mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_flr_tcs.py tests/test_gpu_train.py -q 2>&1 | tail -3
timeout 900 python bench.py --workload train --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_train_tc.json 2> gpurun_out/r02_train_tc.err; echo "train rc=$?"
python - <<'PY'
import json
b=[json.loads(l) for l in open('gpurun_out/r02_train_tc.json') if l.startswith('{')][-1]
print('train: %.1f slices/s, %.1f ms/step, losses %s' % (b['value'],b['ms_per_step'],b['config']['first_losses']))
for k,v in b['rooflines'].items():
    if v: print('   ',k, round(v['ms_per_step'],2),'ms', round(v['frac'],3))
PY
