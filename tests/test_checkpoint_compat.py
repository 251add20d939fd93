"""CPU tier: reference-format checkpoints (models/base_model.py:144-199).  A state_dict saved by the REFERENCE
generator loads strictly into this package's generator and the other way round, the file round trip is bit-exact, and the
file-name convention is the reference's.  (The reference tree is imported from the staged copy or /root/reference; the test is
skipped where neither exists -- the GPU box only needs the round trip.)"""
import importlib
import os
import sys

import pytest
import torch

from conftest import ROOT

TINY = dict(z_dim=64, c_dim=1, w_dim=64, img_resolution=32, mapping_layers=3, channel_base=512, channel_max=48, num_layers=6,
            skip_resolution=16)


def _reference_generator():
    ref = next((d for d in (os.path.join(ROOT, 'baseline', '_ref', 'AFCM'), '/root/reference')
                if os.path.isdir(os.path.join(d, 'models', 'networks', 'stylegan3'))), None)
    if ref is None:
        pytest.skip('reference tree not available')
    sys.dont_write_bytecode = True
    if ref not in sys.path:
        sys.path.insert(0, ref)
    net = importlib.import_module('models.networks.stylegan3.networks_stylegan3')
    torch.manual_seed(3)
    return net.Stylegan3Generator(
        z_dim=64, c_dim=1, w_dim=64, img_resolution=32, img_channels_in=4, img_channels_out=1, mapping_kwargs=dict(num_layers=3),
        synthesis_kwargs=dict(channel_base=512, channel_max=48, num_layers=6, num_critical=2, first_cutoff=2, first_stopband=2 ** 2.1,
                              last_stopband_rel=2 ** 0.3, margin_size=10, output_scale=0.25, skip_resolution=16, conv_kernel=3,
                              filter_size=6, lrelu_upsampling=2, use_radial_filters=False, conv_clamp=256, cond_mod=True))


def test_filename_convention():
    from afcm_b200.checkpoint import network_filename
    assert network_filename('latest', 'G_ema') == 'latest_net_G_ema.pth' and network_filename(40, 'G') == '40_net_G.pth'


def test_round_trip_is_bit_exact(tmp_path):
    from afcm_b200.checkpoint import load_network, save_network
    from afcm_b200.networks_stylegan3 import afcm_generator
    G = afcm_generator(seed=5, device=None, **TINY)
    path = save_network(G, str(tmp_path), 'latest', 'G_ema')
    assert os.path.basename(path) == 'latest_net_G_ema.pth'
    G2 = afcm_generator(seed=6, device=None, **TINY)
    res = load_network(G2, str(tmp_path), 'latest', 'G_ema')
    assert not res.missing_keys and not res.unexpected_keys
    a, b = G.state_dict(), G2.state_dict()
    assert list(a) == list(b) and all(torch.equal(a[k], b[k]) for k in a)
    # a file saved from a DataParallel wrapper carries the `module.` prefix
    torch.save({'module.' + k: v for k, v in a.items()}, str(tmp_path / 'dp.pth'))
    G3 = afcm_generator(seed=7, device=None, **TINY)
    load_network(G3, str(tmp_path / 'dp.pth'))
    assert all(torch.equal(a[k], v) for k, v in G3.state_dict().items())


def test_reference_checkpoint_loads_strictly_both_ways(tmp_path):
    from afcm_b200.checkpoint import load_network, save_network
    from afcm_b200.networks_stylegan3 import afcm_generator
    R = _reference_generator()
    torch.save(R.cpu().state_dict(), str(tmp_path / '7_net_G.pth'))          # models/base_model.py:160
    G = afcm_generator(seed=1, device=None, **TINY)
    res = load_network(G, str(tmp_path), 7, 'G', strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    ra, ga = R.state_dict(), G.state_dict()
    assert list(ra) == list(ga)                                              # same names in the same order
    assert all(ra[k].shape == ga[k].shape and ra[k].dtype == ga[k].dtype and torch.equal(ra[k], ga[k]) for k in ra)
    # and back: a file written here loads into the reference module with its own strict load_state_dict (base_model.py:195)
    G2 = afcm_generator(seed=2, device=None, **TINY)
    path = save_network(G2, str(tmp_path), 'latest', 'G_ema')
    R.load_state_dict(torch.load(path, map_location='cpu', weights_only=True))
    assert all(torch.equal(v, G2.state_dict()[k]) for k, v in R.state_dict().items())
