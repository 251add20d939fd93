#!/usr/bin/env python
"""Small driver for ncu: one exact-fp32 filtered_lrelu forward (sign write) and backward (sign read) at an AFCM layer shape."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from afcm_b200.networks_stylegan3 import design_lowpass_filter
from afcm_b200.torch_utils.ops import filtered_lrelu

dev = torch.device('cuda:0')
if len(sys.argv) > 1: filtered_lrelu.set_train_impl(sys.argv[1])
fu = design_lowpass_filter(12, 64.0, 30.0, 512).to(dev)
fd = design_lowpass_filter(12, 64.0, 30.0, 512).to(dev)
x = torch.randn(8, 64, 278, 278, device=dev, requires_grad=True)
b = torch.zeros(64, device=dev, requires_grad=True)
for _ in range(2):
    y = filtered_lrelu.filtered_lrelu(x, fu=fu, fd=fd, b=b, up=2, down=2, padding=[9, 8, 9, 8], gain=2 ** 0.5, slope=0.2, clamp=256)
    y.sum().backward()
torch.cuda.synchronize()
print('ok', y.shape)
