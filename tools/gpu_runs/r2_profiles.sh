#!/bin/bash
# round-2 final-state profiles on one GPU: bench line, reference arm, per-layer benches, launch list with DRAM bytes, full captures
mkdir -p gpurun_out
S=gpurun_out/r02_summary.txt; : > $S
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench_n1 rc=$?" >> $S
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_ref.err; echo "bench_ref rc=$?" >> $S
timeout 600 python tools/layer_bench.py --batch 64 --ops flrelu_tc,f16in,f16out,nobias --json gpurun_out/r02_layer_bench_flr_tc_fp16.json > gpurun_out/r02_lb_flr.log 2>&1; echo "lb_flr rc=$?" >> $S
timeout 600 python tools/layer_bench.py --batch 64 --ops conv_tc,conv_nchw,f16in,f16out --json gpurun_out/r02_layer_bench_conv_fp16.json > gpurun_out/r02_lb_conv.log 2>&1; echo "lb_conv rc=$?" >> $S
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-fp32-leg --graph 0 > gpurun_out/r02_ncu_bench.log 2>&1; echo "ncu_launches rc=$?" >> $S
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv2d_tc_kernel -s 40 -c 4 -o gpurun_out/r02_prof_conv_direct -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-fp32-leg --graph 0 > gpurun_out/r02_ncu_conv.log 2>&1; echo "ncu_conv rc=$?" >> $S
timeout 300 ncu --set full --clock-control none --import-source on -k regex:flr_tc_kernel -s 1 -c 7 -o gpurun_out/r02_prof_flr_tc -f python tools/flr_tc_prof.py > gpurun_out/r02_ncu_flr_tc.log 2>&1; echo "ncu_flr_tc rc=$?" >> $S
timeout 300 ncu --set full --clock-control none --import-source on -k regex:flr_tc_kernel -s 2 -c 4 -o gpurun_out/r02_prof_flr_tc_signs -f python tools/flr_prof.py tc > gpurun_out/r02_ncu_flr_signs.log 2>&1; echo "ncu_flr_signs rc=$?" >> $S
timeout 300 ncu --set full --clock-control none --import-source on -k regex:flr_t5_kernel -s 1 -c 2 -o gpurun_out/r02_prof_flr_t5 -f python tools/flr_t5_trace.py --size 278 --planes 1024 > gpurun_out/r02_ncu_flr_t5.log 2>&1; echo "ncu_flr_t5 rc=$?" >> $S
for n in conv_direct flr_tc flr_tc_signs flr_t5; do
  python tools/summarize_profiles.py rep gpurun_out/r02_prof_$n.ncu-rep gpurun_out/r02_ncu_$n.txt > /dev/null 2>&1; echo "summary $n rc=$?" >> $S
  rm -f gpurun_out/r02_prof_$n.ncu-rep
done
python tools/summarize_profiles.py launches gpurun_out/r02_launches.csv gpurun_out/r02_launches_final.txt > /dev/null 2>&1; echo "launch summary rc=$?" >> $S
cat $S; cut -c1-400 gpurun_out/r02_bench_n1.json; ls -la gpurun_out | grep r02_
