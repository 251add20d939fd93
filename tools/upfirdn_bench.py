#!/usr/bin/env python
"""Timing of the separable upfirdn2d calls of the training step: the 61-tap blur of the loss images (models/stylegan3_model.py:28,
102-103: filter2d on [B, C, 256, 256]) and the [1,3,3,1] down-sampling of the discriminator (CM/generator.py:664-690), CUDA events,
algorithmic bytes = read x once + write y once per call of the pair of 1-D passes."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from afcm_b200.torch_utils.ops import upfirdn2d  # noqa: E402


def t(fn, n=10):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


dev = torch.device('cuda:0')
sigma = 10.0
size = int(np.floor(sigma * 3))
f61 = torch.arange(-size, size + 1, device=dev).div(sigma).square().neg().exp2()          # stylegan3_model.py:26-28
x = torch.randn(32, 5, 256, 256, device=dev)
ms = t(lambda: upfirdn2d.filter2d(x, f61 / f61.sum()))
print('blur %d taps on %s: %.3f ms, %.0f GB/s (x + y)' % (f61.numel(), tuple(x.shape), ms, 2 * x.numel() * 4 / ms / 1e6))
f4 = upfirdn2d.setup_filter([1, 3, 3, 1], device=dev)
for shape in [(32, 64, 256, 256), (32, 128, 128, 128), (32, 512, 32, 32)]:
    x = torch.randn(*shape, device=dev)
    ms = t(lambda: upfirdn2d.downsample2d(x, f4))
    print('downsample2d [1,3,3,1] on %s: %.3f ms, %.0f GB/s (x + y)' % (shape, ms, (x.numel() * 1.25) * 4 / ms / 1e6))
