# multi-GPU measurements of round 2: training step (NCCL gradient all-reduce) and sharded volume inference on N GPUs of one box
N=${1:-2}
mkdir -p gpurun_out
export NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT NCCL_DEBUG_FILE=gpurun_out/r02_nccl_n${N}_%p.log
if [ "$N" = "1" ]; then
  RUN="python"
else
  RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531"
fi
timeout 900 $RUN bench.py --workload train --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_train_n$N.json 2> gpurun_out/r02_bench_train_n$N.err; echo "train rc=$?"
cat gpurun_out/r02_nccl_n${N}_*.log | grep -hE "NCCL INFO (comm|ncclCommInitRank).*nranks.*COMPLETE|NVLS multicast|Connected all (rings|trees)|Connected NVLS" | sort | uniq -c | sort -rn | head -14 > gpurun_out/r02_nccl_n$N.txt; rm -f gpurun_out/r02_nccl_n${N}_*.log
for T in ${VOLUME_T:-5 2.5}; do
  timeout 600 $RUN tools/volume_bench.py --slices 256 --thickness $T --batch 32 > gpurun_out/r02_volume_n${N}_t$T.json 2> gpurun_out/r02_volume_n${N}_t$T.err; echo "volume t=$T rc=$?"
done
python - <<PY
import json
N="$N"
try:
    b=[json.loads(l) for l in open('gpurun_out/r02_bench_train_n%s.json'%N) if l.startswith('{')][-1]
    print('train N=%s: %.1f slices/s, %.1f ms/step, e2e %.1f' % (N,b['value'],b['ms_per_step'],b['e2e']['value']))
    for k,v in b['rooflines'].items():
        if v: print('   ',k, round(v['ms_per_step'],2),'ms', round(v['frac'],3))
except Exception as e:
    print('train parse failed', e); print(open('gpurun_out/r02_bench_train_n%s.err'%N).read()[-1500:])
for T in ('5','2.5'):
    try:
        v=[json.loads(l) for l in open('gpurun_out/r02_volume_n%s_t%s.json'%(N,T)) if l.startswith('{')][-1]
        print('volume N=%s t=%s: %.1f slices/s, %.2f volumes/s' % (N,T,v['value'],v['volumes_per_sec']))
    except Exception as e:
        print('volume parse failed', T, e); print(open('gpurun_out/r02_volume_n%s_t%s.err'%(N,T)).read()[-800:])
PY
cat gpurun_out/r02_nccl_n$N.txt | cut -c1-200 | head -5
