"""The oracle's torch restatement differentiated by autograd reproduces the gradients the REFERENCE's own autograd
produced in the authoring container (tests/golden/gen_golden.py: modconv.*.dx/dw/ds and tiny_gen_grads.npz), so it can
serve as the checker of the training-step path (BASELINE config 5)."""
import numpy as np
import torch

from conftest import rel_err
from oracle import afcm_oracle as orc

TINY = dict(z_dim=64, c_dim=1, w_dim=64, img_resolution=32, img_channels_in=4, img_channels_out=1,
            mapping_layers=3, channel_base=512, channel_max=48, num_layers=6, skip_resolution=16)


def test_modulated_conv2d_grads(golden_ops):
    g = golden_ops
    for name in ('demod3', 'torgb1', 'demod3b'):
        t = 'modconv.' + name
        demod, pad, ig = int(g[t + '.cfg'][0]), int(g[t + '.cfg'][1]), float(g[t + '.cfg'][2])
        x = torch.as_tensor(g[t + '.x']).requires_grad_(True)
        w = torch.as_tensor(g[t + '.w']).requires_grad_(True)
        s = torch.as_tensor(g[t + '.s']).requires_grad_(True)
        y = orc.t_modulated_conv2d(x, w, s, bool(demod), pad, torch.tensor(ig))
        (y * torch.as_tensor(g[t + '.r'])).sum().backward()
        assert rel_err(x.grad.numpy(), g[t + '.dx']) < 1e-5, name
        assert rel_err(w.grad.numpy(), g[t + '.dw']) < 1e-5, name
        assert rel_err(s.grad.numpy(), g[t + '.ds']) < 1e-5, name


def test_tiny_generator_training_grads(golden_tiny, golden_tiny_grads):
    g, gg = golden_tiny, golden_tiny_grads
    P = {k[2:]: torch.as_tensor(g[k]).clone() for k in g.files if k.startswith('P.')}
    names = [k[2:] for k in gg.files if k.startswith('G.')]
    for n in names:
        P[n].requires_grad_(True)
    y = orc.generator_forward(P, torch.as_tensor(g['z']), torch.as_tensor(g['c']), torch.as_tensor(g['x']), TINY, grad=True)
    loss = (y - torch.as_tensor(gg['target'])).abs().mean()
    assert abs(loss.item() - float(gg['loss'])) < 1e-5
    loss.backward()
    for n in names:
        ref = gg['G.' + n]
        got = P[n].grad.numpy() if P[n].grad is not None else np.zeros_like(ref)
        if np.abs(ref).max() == 0:
            assert np.abs(got).max() < 1e-9, n
        else:
            assert rel_err(got, ref) < 2e-4, n
