mkdir -p gpurun_out
OPS=flrelu_tc,f16in,f16out,nobias
LAYERS=enc4,enc7,enc9,enc11
for i in 1 2; do
timeout 200 python tools/layer_bench.py --batch 64 --ops $OPS --layers $LAYERS | grep -o '"layer": "[a-z0-9]*"\|"flrelu_tc_ms": [0-9.]*' | paste - - | head -4
AFCM_B200_LIB=$PWD/afcm_b200/libafcm_b200_b3.so timeout 200 python tools/layer_bench.py --batch 64 --ops $OPS --layers $LAYERS | grep -o '"layer": "[a-z0-9]*"\|"flrelu_tc_ms": [0-9.]*' | paste - - | head -4 | sed 's/^/B3 /'
done
