mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_generator.py tests/test_gpu_reference_network.py tests/test_gpu_tc.py -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-fp32-leg > gpurun_out/r2_bench_q.json 2> gpurun_out/r2_bench_q.err; echo "bench rc=$?"
timeout 600 python bench.py --steps 50 --warmup 5 --batch 1 --no-cpu-baseline --no-fp32-leg > gpurun_out/r2_bench_b1.json 2> gpurun_out/r2_bench_b1.err; echo "bench b1 rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/r2_bench_q.json','gpurun_out/r2_bench_b1.json'):
    try:
        b=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(b['value'],1),'slices/s', round(b['ms_per_step'],3),'ms/step launches',b['gpu_launches'], 'e2e', round(b['e2e']['value'],1), b['clocks']['sm_mhz'])
    except Exception as e:
        print('parse failed', f, e); print(open(f.replace('.json','.err')).read()[-800:])
PY
