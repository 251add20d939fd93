# round-2 validation: GPU test tier, smoke, forward bench (one JSON line)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -v "^$" > gpurun_out/r2_pytest_gpu.log; echo "pytest rc=${PIPESTATUS[0]}"
tail -5 gpurun_out/r2_pytest_gpu.log
grep -E "fast path|flr_tc fast variant|sign tensor" gpurun_out/r2_pytest_gpu.log | head -80
timeout 600 python __graft_entry__.py smoke > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    b=json.loads(open('gpurun_out/r2_bench.json').read().strip().splitlines()[-1])
    print({k:b[k] for k in ('value','ms_per_step','gpu_launches','parity','fp32_path','clocks')})
    print('e2e',b['e2e'])
    for k,v in b['rooflines'].items():
        if v: print(k, round(v['ms_per_step'],2),'ms', round(v['frac'],3))
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/r2_bench.err').read()[-2000:])
PY
