#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/summary.txt; : > $S
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_generator.py -m gpu -q -x > gpurun_out/t_gpu.log 2>&1; echo "pytest_gpu rc=$?" >> $S
timeout 600 python tools/layer_bench.py --batch 16 --ops conv_tc,f16in,f16out --json gpurun_out/lb_conv16.json > gpurun_out/lb_conv16.log 2>&1; echo "lb_conv16 rc=$?" >> $S
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> $S
cat $S; tail -4 gpurun_out/t_gpu.log; cat gpurun_out/bench.log | cut -c1-300; grep SUMMARY gpurun_out/lb_conv16.log
python - <<'PY'
import json
for r in json.load(open('gpurun_out/lb_conv16.json'))['rows']:
    if 'conv_tc_ms' in r: print(r['layer'], r['cin'], r['cout'], r['H'], round(r['conv_tc_ms'],3), round(r['conv_tc_tflops']))
PY
