#!/usr/bin/env python
"""Forward (sign write) + backward (sign read) time of filtered_lrelu per AFCM layer geometry at batch 32: exact fp32 kernel
vs tensor-core kernel (flr_tcs)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from afcm_b200.networks_stylegan3 import design_lowpass_filter
from afcm_b200.torch_utils.ops import filtered_lrelu

dev = torch.device('cuda:0')
f12 = design_lowpass_filter(12, 64.0, 30.0, 512).to(dev)
f24 = design_lowpass_filter(24, 32.0, 30.0, 512).to(dev)
# (name, C, H, up, down, fu, fd, padding, count in the generator)
CASES = [('u2d2 278->276 c64', 64, 278, 2, 2, f12, f12, [9, 8, 9, 8], 4), ('u2d2 278->276 c128', 128, 278, 2, 2, f12, f12, [9, 8, 9, 8], 2),
         ('u2d4 278->148 c181', 181, 278, 2, 4, f12, f24, [34, 33, 34, 33], 1), ('u2d2 150->148 c256', 256, 150, 2, 2, f12, f12, [9, 8, 9, 8], 4),
         ('u2d4 150->84 c512', 512, 150, 2, 4, f12, f24, [34, 33, 34, 33], 1), ('u2d2 86->84 c512', 512, 86, 2, 2, f12, f12, [9, 8, 9, 8], 2),
         ('u4d2 86->148 c362', 362, 86, 4, 2, f24, f12, [-6, -9, -6, -9], 1), ('u4d2 150->276 c128', 128, 150, 4, 2, f24, f12, [-6, -9, -6, -9], 1),
         ('u2d2 38->36 c512', 512, 38, 2, 2, f12, f12, [9, 8, 9, 8], 6)]


def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


N = 32
for name, C, H, up, dn, fu, fd, pad, cnt in CASES:
    x = torch.randn(N, C, H, H, device=dev, requires_grad=True)
    b = torch.zeros(C, device=dev, requires_grad=True)
    row = [name]
    for impl in ('exact', 'tcs', 'tc'):
        filtered_lrelu.set_train_impl(impl)
        y = filtered_lrelu.filtered_lrelu(x, fu=fu, fd=fd, b=b, up=up, down=dn, padding=pad, gain=2 ** 0.5, slope=0.2, clamp=256)
        g = torch.randn_like(y)
        tf = timeit(lambda: filtered_lrelu.filtered_lrelu(x, fu=fu, fd=fd, b=b, up=up, down=dn, padding=pad, gain=2 ** 0.5, slope=0.2, clamp=256))
        tb = timeit(lambda: torch.autograd.grad(y, x, g, retain_graph=True))
        gb = 4.0 * (x.numel() + y.numel()) / 1e9
        row.append('%s fwd %6.2f ms (%5.0f GB/s) bwd %6.2f ms' % (impl, tf, gb / tf * 1e3, tb))
        del y, g
    print('  |  '.join(row))
    del x
