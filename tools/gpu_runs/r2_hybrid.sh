#!/bin/bash
# hybrid pack / direct dispatch of the tcgen05 convolution: parity tests, then the forward bench with and without it
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_generator.py -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 --no-fp32-leg > gpurun_out/bench_hybrid.json 2> gpurun_out/bench_hybrid.err; echo "bench rc=$?"
AFCM_HYBRID_PACK=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-fp32-leg > gpurun_out/bench_nohybrid.json 2> gpurun_out/bench_nohybrid.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/bench_hybrid.json','gpurun_out/bench_nohybrid.json'):
    try:
        b=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(b['value'],1), round(b['ms_per_step'],2), b['gpu_launches'], b['parity']['rel_err'], b['clocks']['sm_mhz'])
        for k,v in b['rooflines'].items():
            if v: print('   ',k, round(v['ms_per_step'],2),'ms', round(v['frac'],3), v['launches_per_step'])
    except Exception as e:
        print(f, 'parse failed', e)
PY
