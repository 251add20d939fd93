mkdir -p gpurun_out
for prec in fp32 tc; do
timeout 900 python bench.py --workload train --steps 1 --warmup 3 --no-cpu-baseline --precision $prec > gpurun_out/r02_train_p$prec.json 2> gpurun_out/r02_train_p$prec.err; echo "train $prec rc=$?"
done
python - <<'PY'
import json
for f in ('fp32','tc'):
    try:
        b=[json.loads(l) for l in open('gpurun_out/r02_train_p%s.json'%f) if l.startswith('{')][-1]
        print(f,'train: %.1f slices/s, %.1f ms/step, losses %s' % (b['value'],b['ms_per_step'],b['config']['first_losses']))
    except Exception as e:
        print('parse failed', f, e); print(open('gpurun_out/r02_train_p%s.err'%f).read()[-800:])
PY
