"""Boundary proof: the REFERENCE's own generator code -- models/networks/stylegan3/
networks_stylegan3.py, unmodified, imported from the staged copy of the reference tree -- runs on this library's operators:
its `filtered_lrelu`, `bias_act`, `conv2d_gradfix` module references and `modulated_conv2d` (call sites NET:365, 371, 505, 510,
100) and the `conv2d_resample` / `bias_act` references of CoModGAN/layers.py (Conv2dLayer e_16x16, CM/layers.py:157,161) are
pointed at afcm_b200.torch_utils.ops, nothing else changes.  The result must equal the golden output that the same code
produced with the reference's own CPU `_ref` operators (tests/golden/full_gen.npz, tiny_gen.npz).

The reference tree is test infrastructure: it is staged by tools/stage_reference.py under baseline/_ref/AFCM (git-ignored, it
travels to the GPU box); the test is skipped when it is absent."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, rel_err

pytestmark = pytest.mark.gpu

REF_DIRS = [os.path.join(ROOT, 'baseline', '_ref', 'AFCM'), '/root/reference']


def _reference_modules():
    ref = next((d for d in REF_DIRS if os.path.isdir(os.path.join(d, 'models', 'networks', 'stylegan3'))), None)
    if ref is None:
        pytest.skip('reference tree not staged (python tools/stage_reference.py)')
    sys.dont_write_bytecode = True
    if ref not in sys.path:
        sys.path.insert(0, ref)
    try:
        net = importlib.import_module('models.networks.stylegan3.networks_stylegan3')
        cml = importlib.import_module('models.networks.CoModGAN.layers')
    except Exception as e:                                   # a dependency of the reference missing on this machine
        pytest.skip(f'reference modules do not import here: {e!r}')
    return net, cml


class _ConvResample:
    """conv2d_resample module stand-in for Conv2dLayer (CM/layers.py:157): up = down = 1 is what AFCM instantiates."""

    @staticmethod
    def conv2d_resample(x, w, f=None, up=1, down=1, padding=0, groups=1, flip_weight=True, flip_filter=False):
        from afcm_b200.torch_utils.ops import conv2d_gradfix
        assert up == 1 and down == 1 and groups == 1 and flip_weight
        return conv2d_gradfix.conv2d(x, w, padding=padding)


def _swap_ops(monkeypatch, net, cml):
    from afcm_b200 import networks_stylegan3 as ours
    from afcm_b200.torch_utils.ops import bias_act, conv2d_gradfix, filtered_lrelu
    monkeypatch.setattr(net, 'filtered_lrelu', filtered_lrelu)
    monkeypatch.setattr(net, 'bias_act', bias_act)
    monkeypatch.setattr(net, 'conv2d_gradfix', conv2d_gradfix)
    monkeypatch.setattr(net, 'modulated_conv2d', ours.modulated_conv2d)
    monkeypatch.setattr(cml, 'conv2d_resample', _ConvResample)
    monkeypatch.setattr(cml, 'bias_act', bias_act)


def _build_ref(net, cfg, seed):
    torch.manual_seed(seed)
    return net.Stylegan3Generator(
        z_dim=cfg['z_dim'], c_dim=1, w_dim=cfg['w_dim'], img_resolution=cfg['img_resolution'], img_channels_in=4, img_channels_out=1,
        mapping_kwargs=dict(num_layers=cfg['mapping_layers']),
        synthesis_kwargs=dict(channel_base=cfg['channel_base'], channel_max=cfg['channel_max'], num_layers=cfg['num_layers'],
                              num_critical=2, first_cutoff=2, first_stopband=2 ** 2.1, last_stopband_rel=2 ** 0.3, margin_size=10,
                              output_scale=0.25, skip_resolution=cfg['skip_resolution'], conv_kernel=3, filter_size=6,
                              lrelu_upsampling=2, use_radial_filters=False, conv_clamp=256,
                              magnitude_ema_beta=0.5 ** (16 / (20 * 1e3)), cond_mod=True)).eval()


TINY = dict(z_dim=64, w_dim=64, img_resolution=32, mapping_layers=3, channel_base=512, channel_max=48, num_layers=6, skip_resolution=16)
FULL = dict(z_dim=512, w_dim=512, img_resolution=256, mapping_layers=8, channel_base=16384, channel_max=512, num_layers=14,
            skip_resolution=128)


def test_reference_generator_runs_unmodified_on_the_swapped_ops_tiny(monkeypatch, golden_tiny):
    net, cml = _reference_modules()
    g = golden_tiny
    dev = torch.device('cuda:0')
    G = _build_ref(net, TINY, seed=0)
    sd = {k[2:]: torch.as_tensor(g[k]) for k in g.files if k.startswith('P.')}
    G.load_state_dict(sd, strict=False)
    G = G.to(dev)
    _swap_ops(monkeypatch, net, cml)
    from afcm_b200 import _lib
    n0 = _lib.launch_count()
    with torch.no_grad():
        y = G(torch.as_tensor(g['z'], device=dev), torch.as_tensor(g['c'], device=dev), torch.as_tensor(g['x'], device=dev),
              noise_mode='const')
    assert _lib.launch_count() - n0 > 30                    # the native kernels ran (no reference plugin, no torch conv)
    assert rel_err(y.cpu().numpy(), g['y']) < 1e-4


def test_reference_generator_runs_unmodified_on_the_swapped_ops_full(monkeypatch, golden_full):
    """The full 58.5 M-parameter network of BASELINE configs 1/2 (seeded init of the reference itself), B = 2, fp32 path."""
    net, cml = _reference_modules()
    g = golden_full
    dev = torch.device('cuda:0')
    G = _build_ref(net, FULL, seed=0).to(dev)
    _swap_ops(monkeypatch, net, cml)
    x = (torch.as_tensor(g['x_u8']).float() * (2.0 / 255.0) - 1.0).clamp(-1, 1).to(dev)
    with torch.no_grad():
        y = G(torch.as_tensor(g['z'], device=dev), torch.as_tensor(g['c'], device=dev), x, noise_mode='const')
    assert rel_err(y.cpu().numpy(), g['y']) < 1e-4
    # and on the tensor-core convolution path (fp16 operands, fp32 storage): the documented 16-bit bound
    from afcm_b200 import inference
    inference.set_precision('tc')
    try:
        with torch.no_grad():
            y16 = G(torch.as_tensor(g['z'], device=dev), torch.as_tensor(g['c'], device=dev), x, noise_mode='const')
    finally:
        inference.set_precision('fp32')
    assert rel_err(y16.cpu().numpy(), g['y']) < 6e-3
