"""fma -- fused multiply-add a * b + c with broadcasting (reference: models/networks/CoModGAN/torch_utils/ops/fma.py; the
reference's version exists for its custom double-backward).  Used by CM/layers.py:62 (demodulation + noise) off the AFCM
generator path; evaluated with tensor arithmetic (first and higher-order gradients come from autograd)."""
import torch


def fma(a, b, c):
    return torch.addcmul(c, a, b)
