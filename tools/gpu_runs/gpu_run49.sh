#!/bin/bash
# final state of round 1 (one GPU): full GPU tier, smoke, bench (both arms), training bench, volume bench, ncu launch list with DRAM bytes
mkdir -p gpurun_out
S=gpurun_out/summary49.txt; : > $S
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 >> $S
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" >> $S 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.log 2> gpurun_out/bench_n1.err; echo "bench_n1 rc=$?" >> $S
timeout 600 python bench.py --workload train --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_train_n1.log 2> gpurun_out/bench_train_n1.err; echo "bench_train rc=$?" >> $S
timeout 300 python tools/volume_bench.py > gpurun_out/volume_bench.log 2> gpurun_out/volume_bench.err; echo "volume rc=$?" >> $S
timeout 300 python tools/layer_bench.py --batch 64 --ops flrelu_tc,f16in,f16out,nobias --json gpurun_out/flr_final.json > gpurun_out/flr_final.log 2>&1; echo "flr layer rc=$?" >> $S
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --graph 0 > gpurun_out/ncu_bench.log 2>&1; echo "ncu_launches rc=$?" >> $S
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "bench_ref rc=$?" >> $S
cat $S; cut -c1-250 gpurun_out/bench_n1.log; cut -c1-200 gpurun_out/bench_train_n1.log; cut -c1-200 gpurun_out/volume_bench.log; grep SUMMARY gpurun_out/flr_final.log | cut -c1-400
