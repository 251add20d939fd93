#!/usr/bin/env python
"""Small driver for ncu: the tcgen05 weight gradient at three AFCM layer shapes (batch 32)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from afcm_b200 import _lib
from afcm_b200.torch_utils.ops import conv2d_gradfix as cg
dev = torch.device('cuda:0'); L = _lib.lib(); dt = torch.bfloat16; code = _lib.dtype_code(dt); st = _lib.stream_ptr(dev); N = 32
for Ci, Co, H in ((362, 512, 148), (512, 512, 84), (128, 181, 276)):
    x = torch.randn(N, Ci, H, H, device=dev); dy = torch.randn(N, Co, H + 2, H + 2, device=dev)
    xp, dyp = cg._pack(x, None, dt), cg._pack(dy, None, dt)
    dw = torch.empty(Co, Ci, 3, 3, device=dev)
    nbytes = int(L.afcm_conv2d_wgrad_tc_workspace(N, Ci, H, H, Co, 2)); ws = torch.empty(nbytes // 4, device=dev)
    for _ in range(2):
        _lib.check(L.afcm_conv2d_wgrad_tc5(_lib.ptr(dyp), _lib.ptr(xp), _lib.ptr(dw), _lib.ptr(ws), nbytes, code, N, Ci, H, H, Co, 2, st))
    torch.cuda.synchronize()
print('ok')
