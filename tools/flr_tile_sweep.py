#!/usr/bin/env python
"""Tile sweep of the exact-fp32 filtered_lrelu (forward with sign write, backward with sign read) at AFCM layer shapes."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from afcm_b200 import _lib
from afcm_b200.networks_stylegan3 import design_lowpass_filter
from afcm_b200.torch_utils.ops import filtered_lrelu

dev = torch.device('cuda:0')
L = _lib.lib()
f12 = design_lowpass_filter(12, 64.0, 30.0, 512).to(dev)
f24 = design_lowpass_filter(24, 32.0, 30.0, 512).to(dev)
CASES = [('u2d2 278->276', 64, 278, 2, 2, f12, f12, [9, 8, 9, 8]), ('u2d4 278->148', 181 // 4, 278, 2, 4, f12, f24, [34, 33, 34, 33]),
         ('u4d2 150->276', 32, 150, 4, 2, f24, f12, [-6, -9, -6, -9]), ('u2d2 38->36', 512, 38, 2, 2, f12, f12, [9, 8, 9, 8])]
TILES = [(0, 0), (96, 32), (96, 16), (64, 64), (64, 48), (64, 32), (64, 24), (64, 16), (48, 48), (48, 32), (48, 24), (40, 40), (32, 32),
         (32, 24), (32, 16), (24, 24), (16, 16)]


def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for name, C, H, up, dn, fu, fd, pad in CASES:
    x = torch.randn(8, C, H, H, device=dev, requires_grad=True)
    b = torch.zeros(C, device=dev, requires_grad=True)
    res = []
    for tw, th in TILES:
        L.afcm_filtered_lrelu_set_tile(tw, th)
        try:
            y = filtered_lrelu.filtered_lrelu(x, fu=fu, fd=fd, b=b, up=up, down=dn, padding=pad, gain=2 ** 0.5, slope=0.2, clamp=256)
            g = torch.randn_like(y)
            tf = timeit(lambda: filtered_lrelu.filtered_lrelu(x, fu=fu, fd=fd, b=b, up=up, down=dn, padding=pad, gain=2 ** 0.5, slope=0.2, clamp=256))
            tb = timeit(lambda: torch.autograd.grad(y, x, g, retain_graph=True))
            res.append((tf + tb, tf, tb, tw, th))
        except Exception as e:
            res.append((float('inf'), 0, 0, tw, th))
    L.afcm_filtered_lrelu_set_tile(0, 0)
    auto = [r for r in res if r[3] == 0][0]
    print(name, ' auto: fwd %.3f bwd %.3f ms' % (auto[1], auto[2]))
    for r in sorted(res)[:6]:
        print('    tile %3dx%-3d fwd %.3f bwd %.3f total %.3f' % (r[3], r[4], r[1], r[2], r[0]))
    bf = min(res, key=lambda r: r[1] if r[1] > 0 else 1e9); bb = min(res, key=lambda r: r[2] if r[2] > 0 else 1e9)
    print('    best fwd tile %dx%d %.3f; best bwd tile %dx%d %.3f' % (bf[3], bf[4], bf[1], bb[3], bb[4], bb[2]))
