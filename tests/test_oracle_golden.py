"""Pins the CPU oracle (oracle/) against golden vectors produced by the reference's own `_ref`
implementation (tests/golden/gen_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import afcm_oracle as orc

TOL = 2e-6   # float32 rounding; summation order differs from torch's conv kernels


def _cfg(g, k):
    c = g[k + '.cfg']
    up, dn = int(c[0]), int(c[1]); pad = [int(v) for v in c[2:6]]
    gain, slope, clamp = float(c[6]), float(c[7]), (None if c[8] < 0 else float(c[8]))
    fu = g[k + '.fu']; fd = g[k + '.fd']
    return up, dn, pad, gain, slope, clamp, (fu if fu.size else None), (fd if fd.size else None)


def test_filtered_lrelu_c_oracle(golden_ops):
    g = golden_ops
    for name in g['flrelu.names']:
        k = 'flrelu.' + str(name)
        up, dn, pad, gain, slope, clamp, fu, fd = _cfg(g, k)
        y = orc.filtered_lrelu(g[k + '.x'], fu, fd, g[k + '.b'], up, dn, pad, gain, slope, clamp)
        assert y.shape == g[k + '.y'].shape, name
        assert rel_err(y, g[k + '.y']) < TOL, name
        yf = orc.filtered_lrelu(g[k + '.x'], fu, fd, g[k + '.b'], up, dn, pad, gain, slope, clamp, flip_filter=True)
        assert rel_err(yf, g[k + '.yflip']) < TOL, name


def test_filtered_lrelu_backward_via_signs(golden_ops):
    """dx of the reference autograd == the op re-applied with swapped filters reading the sign tensor
    (OPS/filtered_lrelu.py:252-266); db = dx.sum([0,2,3])."""
    g = golden_ops
    for name in g['flrelu.names']:
        k = 'flrelu.' + str(name)
        up, dn, pad, gain, slope, clamp, fu, fd = _cfg(g, k)
        x = g[k + '.x']
        y, so = orc.filtered_lrelu(x, fu, fd, g[k + '.b'], up, dn, pad, gain, slope, clamp, write_signs=True)
        fun = 1 if fu is None else fu.shape[0]; fdn = 1 if fd is None else fd.shape[0]
        xh, xw = x.shape[2:]; yh, yw = y.shape[2:]
        pp = [(fun - 1) + (fdn - 1) - pad[0], xw * up - yw * dn + pad[0] - (up - 1),
              (fun - 1) + (fdn - 1) - pad[2], xh * up - yh * dn + pad[2] - (up - 1)]
        dx = orc.filtered_lrelu(g[k + '.r'], fd, fu, None, dn, up, pp, gain * up ** 2 / dn ** 2, slope, None,
                                flip_filter=True, si=so, sx=-(fun - 1) + pad[0], sy=-(fun - 1) + pad[2])
        assert dx.shape == x.shape, name
        assert rel_err(dx, g[k + '.dx']) < 5e-6, name
        assert rel_err(dx.sum((0, 2, 3)), g[k + '.db']) < 2e-5, name


def test_filtered_lrelu_torch_restatement(golden_ops):
    g = golden_ops
    for name in g['flrelu.names']:
        k = 'flrelu.' + str(name)
        up, dn, pad, gain, slope, clamp, fu, fd = _cfg(g, k)
        y = orc.t_filtered_lrelu(torch.as_tensor(g[k + '.x']), fu, fd, torch.as_tensor(g[k + '.b']), up, dn, pad,
                                 gain, slope, clamp)
        assert rel_err(y.numpy(), g[k + '.y']) < TOL, name


def test_upfirdn2d(golden_ops):
    g = golden_ops
    for name in g['upfirdn.names']:
        k = 'upfirdn.' + str(name)
        c = g[k + '.cfg']
        f = g[k + '.f']; f = f if f.size else None
        y = orc.upfirdn2d(g[k + '.x'], f, up=(int(c[0]), int(c[1])), down=(int(c[2]), int(c[3])),
                          padding=[int(v) for v in c[4:8]], flip_filter=bool(c[8]), gain=float(c[9]))
        assert y.shape == g[k + '.y'].shape, name
        assert rel_err(y, g[k + '.y']) < TOL, name


@pytest.mark.parametrize('act', list(orc.ACT_IDX))
def test_bias_act(golden_ops, act):
    g = golden_ops
    for tag, clamp in (('n', None), ('c', 0.7)):
        y = orc.bias_act(g['bias_act.x'], g['bias_act.b'], act=act, clamp=clamp)
        ref = g[f'bias_act.{act}.{tag}.y']
        assert rel_err(y, ref) < 3e-6, (act, tag)
        dy = np.full_like(ref, 0.5)
        dx = orc.bias_act_grad(dy, g['bias_act.b'], g['bias_act.x'], ref, act=act, clamp=clamp)
        assert rel_err(dx, g[f'bias_act.{act}.{tag}.dx']) < 3e-5, (act, tag)
    y2 = orc.bias_act(g['bias_act.x2'], g['bias_act.b2'], act='lrelu')
    assert rel_err(y2, g['bias_act.y2']) < TOL


def test_modulated_conv2d(golden_ops):
    g = golden_ops
    for name in ('demod3', 'torgb1', 'demod3b'):
        t = 'modconv.' + name
        demod, pad, ig = int(g[t + '.cfg'][0]), int(g[t + '.cfg'][1]), float(g[t + '.cfg'][2])
        y = orc.modulated_conv2d(g[t + '.x'], g[t + '.w'], g[t + '.s'], bool(demod), pad, np.float32(ig))
        assert rel_err(y, g[t + '.y']) < 3e-6, name
        yt = orc.t_modulated_conv2d(torch.as_tensor(g[t + '.x']), torch.as_tensor(g[t + '.w']),
                                    torch.as_tensor(g[t + '.s']), bool(demod), pad, torch.tensor(ig))
        assert rel_err(yt.numpy(), g[t + '.y']) < 3e-6, name


def test_fully_connected(golden_ops):
    g = golden_ops
    y = orc.fully_connected(g['fc.x'], g['fc.w'], g['fc.b'], 0.01 / np.sqrt(37), 0.01, 'lrelu')
    assert rel_err(y, g['fc.y']) < TOL
    y = orc.fully_connected(g['fc.x'], g['fc2.w'], g['fc2.b'], 1 / np.sqrt(37), 1.0, 'linear')
    assert rel_err(y, g['fc2.y']) < TOL
    yt = orc.t_fc(torch.as_tensor(g['fc.x']), torch.as_tensor(g['fc.w']), torch.as_tensor(g['fc.b']), 0.01, 'lrelu')
    assert rel_err(yt.numpy(), g['fc.y']) < TOL


def test_mapping(golden_ops):
    g = golden_ops
    P = {'mapping.' + k[len('map.P.'):]: torch.as_tensor(g[k]) for k in g.files if k.startswith('map.P.')}
    cfg = dict(z_dim=48, c_dim=1, w_dim=40, mapping_layers=3, num_layers=4)
    ws = orc.mapping_forward(P, torch.as_tensor(g['map.z']), torch.as_tensor(g['map.c']), cfg)
    assert ws.shape == g['map.ws'].shape
    assert rel_err(ws.numpy(), g['map.ws']) < 5e-6


TINY = dict(z_dim=64, c_dim=1, w_dim=64, img_resolution=32, img_channels_in=4, img_channels_out=1,
            mapping_layers=3, channel_base=512, channel_max=48, num_layers=6, skip_resolution=16)


def test_tiny_generator(golden_tiny):
    g = golden_tiny
    P = {k[2:]: torch.as_tensor(g[k]) for k in g.files if k.startswith('P.')}
    taps = {}
    y = orc.generator_forward(P, torch.as_tensor(g['z']), torch.as_tensor(g['c']), torch.as_tensor(g['x']), TINY, taps)
    for k, v in taps.items():
        ref = g['tap.' + k]
        got = (v[:1, :8] if v.ndim == 4 else v).numpy()
        assert rel_err(got, ref) < 2e-5, k
    assert rel_err(y.numpy(), g['y']) < 2e-5
    # the filters designed by the oracle equal the reference's registered buffers bit-for-bit
    enc, syn, _, _, _ = orc.layer_specs(TINY)
    for i, sp in enumerate(enc):
        assert np.array_equal(sp['up_filter'], g[f'F.synthesis.encoder_{i}.up_filter'])
        assert np.array_equal(sp['down_filter'], g[f'F.synthesis.encoder_{i}.down_filter'])


def test_full_generator_seeded_init_matches_reference(golden_full):
    """oracle.init_params(seed=0) reproduces the reference's torch.manual_seed(0) random init."""
    g = golden_full
    P = orc.init_params(seed=0)
    keys = [k for k in g.files if k.startswith('S.') and not k.endswith('_filter') and 'w_avg' not in k]
    assert len(keys) > 100
    for k in keys:
        v = P[k[2:]].double().flatten()
        got = np.asarray([v.sum().item(), v.abs().sum().item(), v[:: max(1, v.numel() // 7)].sum().item()])
        assert np.allclose(got, g[k], rtol=0, atol=0), k
