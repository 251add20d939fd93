#!/usr/bin/env python
"""Small driver for ncu: the tensor-core filtered_lrelu forward (flr_tc_kernel, fast variant: fp16 planes, bias already
added, no skip) at one AFCM layer shape per geometry, batch 64 -- the launches of the benchmarked inference path.

    ncu --set full --clock-control none --import-source on -k regex:flr_tc_kernel -o gpurun_out/prof_flr_tc python tools/flr_tc_prof.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from afcm_b200.networks_stylegan3 import afcm_generator  # noqa: E402
from afcm_b200.torch_utils.ops.filtered_lrelu import filtered_lrelu_tc  # noqa: E402

dev = torch.device('cuda:0')
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
S = afcm_generator(seed=0, device=dev).synthesis
# enc3: 128 ch 278 -> 276 (up 2, down 2); enc4: 181 ch 278 -> 148 (2, 4); L10: 128 ch 150 -> 276 (4, 2); enc12: 512 ch 38 -> 36
picks = [S.encoder_3, S.encoder_4, getattr(S, [n for n in S.layer_names if n.startswith('L10_')][0]), S.encoder_12]
for L in picks:
    C, Hc = L.out_channels, int(L.in_size[0]) + 2
    x = torch.randn(B, C, Hc, Hc, device=dev).half()
    for _ in range(2):
        y = filtered_lrelu_tc(x, L.up_filter, L.down_filter, None, up=L.up_factor, down=L.down_factor, padding=L.padding,
                              gain=2 ** 0.5, slope=0.2, clamp=256.0, out_dtype=torch.float16)
    torch.cuda.synchronize()
    print('ok', L.up_factor, L.down_factor, tuple(x.shape), tuple(y.shape))
    del x, y
