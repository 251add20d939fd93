#!/usr/bin/env python
"""Golden vectors of the CoModGAN networks AFCM instantiates next to the stylegan3 generator: the
discriminator `CoModDiscriminator` (models/networks/CoModGAN/generator.py:781-836; AFCM config: c_dim 1, img_channels 5, resnet
blocks, minibatch-std epilogue) and the baseline generator `CoModGenerator` (:546-575), small instances, run by the REFERENCE's own
code with its CPU `_ref` operators.  Writes tests/golden/cm_nets.npz (parameters, inputs, outputs).

    python tests/golden/gen_golden_cm.py            (in the authoring container: needs /root/reference)
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
sys.path.insert(0, '/root/reference')

D_CFG = dict(c_dim=1, img_resolution=32, img_channels=5, channel_base=256, channel_max=32, epilogue_kwargs=dict(mbstd_group_size=2))
G_CFG = dict(z_dim=32, c_dim=1, w_dim=32, img_resolution=32, img_channels_in=4, img_channels_out=1,
             mapping_kwargs=dict(name='MappingNetwork', num_layers=2),
             synthesis_kwargs=dict(name='SynthesisNetwork', channel_base=256, channel_max=32))


def main():
    from models.networks.CoModGAN.generator import CoModDiscriminator, CoModGenerator
    out = {}
    torch.manual_seed(0)
    D = CoModDiscriminator(**D_CFG).eval()
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for p in D.parameters():                      # non-zero biases
            if p.ndim == 1:
                p.copy_(torch.randn(p.shape, generator=g) * 0.1)
    img = torch.randn(4, 5, 32, 32, generator=g)
    c = torch.rand(4, 1, generator=g)
    with torch.no_grad():
        logits = D(img, c)
    out.update({'D.P.' + k: v.numpy() for k, v in D.state_dict().items()})
    out.update({'D.img': img.numpy(), 'D.c': c.numpy(), 'D.y': logits.numpy()})
    # first-order gradients of the discriminator (the non-saturating logistic loss term of the generator step)
    for p in D.parameters():
        p.requires_grad_(True)
    img_g = img.clone().requires_grad_(True)
    torch.nn.functional.softplus(-D(img_g, c)).mean().backward()
    out['D.dimg'] = img_g.grad.numpy()
    out.update({'D.G.' + k: p.grad.numpy() for k, p in D.named_parameters() if p.grad is not None})
    # R1 regularisation (models/comodgan_model.py:128-161): gradient of the squared input-gradient norm -- a double backward
    for p in D.parameters():
        p.grad = None
    img_r = img.clone().requires_grad_(True)
    r1_grads = torch.autograd.grad(outputs=[D(img_r, c).sum()], inputs=[img_r], create_graph=True, only_inputs=True)[0]
    pen = r1_grads.square().sum([1, 2, 3])
    (pen * (10.0 / 2)).mean().backward()
    out['D.r1_pen'] = pen.detach().numpy()
    out.update({'D.R1.' + k: p.grad.numpy() for k, p in D.named_parameters() if p.grad is not None})

    torch.manual_seed(2)
    G = CoModGenerator(**G_CFG).eval()
    z = torch.randn(2, 32, generator=g); cg = torch.rand(2, 1, generator=g)
    x = torch.rand(2, 4, 32, 32, generator=g) * 2 - 1
    with torch.no_grad():
        y = G(z, cg, x, noise_mode='const')
    out.update({'G.P.' + k: v.numpy() for k, v in G.state_dict().items()})
    out.update({'G.z': z.numpy(), 'G.c': cg.numpy(), 'G.x': x.numpy(), 'G.y': y.numpy()})
    np.savez_compressed(os.path.join(HERE, 'cm_nets.npz'), **out)
    print('cm_nets.npz', len(out), 'arrays; D logits', logits.flatten().tolist(), 'G out', tuple(y.shape), float(y.abs().max()))


if __name__ == '__main__':
    main()
