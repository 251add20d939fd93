"""SURVEY 8(f) rows 3 and 4 through the operator API: the REFERENCE's own CoModGAN networks -- `CoModDiscriminator` (the
discriminator of every AFCM model, models/networks/CoModGAN/generator.py:781-836: resnet blocks of Conv2dLayer with [1,3,3,1]
down-sampling, minibatch-std epilogue, conditioning projection) and the baseline generator `CoModGenerator` (:546-575:
StyleGAN2-style modulated convolutions with up = 2 transposed convolutions, noise inputs, ToRGB up-sampling) -- imported
unmodified from the staged reference tree, run on this library's operators: the `conv2d_resample`, `upfirdn2d`, `bias_act` and
`fma` module references of CM/layers.py and CM/generator.py are pointed at afcm_b200.torch_utils.ops.  Outputs (and the
discriminator's first-order gradients) must equal what the same code produced with the reference's CPU `_ref` operators
(tests/golden/cm_nets.npz, minted by tests/golden/gen_golden_cm.py)."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT, rel_err

pytestmark = pytest.mark.gpu

D_CFG = dict(c_dim=1, img_resolution=32, img_channels=5, channel_base=256, channel_max=32, epilogue_kwargs=dict(mbstd_group_size=2))
G_CFG = dict(z_dim=32, c_dim=1, w_dim=32, img_resolution=32, img_channels_in=4, img_channels_out=1,
             mapping_kwargs=dict(name='MappingNetwork', num_layers=2),
             synthesis_kwargs=dict(name='SynthesisNetwork', channel_base=256, channel_max=32))


@pytest.fixture(scope='module')
def golden_cm():
    return np.load(os.path.join(GOLDEN, 'cm_nets.npz'))


def _reference_cm(monkeypatch):
    ref = next((d for d in (os.path.join(ROOT, 'baseline', '_ref', 'AFCM'), '/root/reference')
                if os.path.isdir(os.path.join(d, 'models', 'networks', 'CoModGAN'))), None)
    if ref is None:
        pytest.skip('reference tree not staged (python tools/stage_reference.py)')
    sys.dont_write_bytecode = True
    if ref not in sys.path:
        sys.path.insert(0, ref)
    try:
        gen = importlib.import_module('models.networks.CoModGAN.generator')
        lay = importlib.import_module('models.networks.CoModGAN.layers')
    except Exception as e:
        pytest.skip(f'reference modules do not import here: {e!r}')
    from afcm_b200.torch_utils.ops import bias_act, conv2d_resample, fma, upfirdn2d
    for mod in (gen, lay):
        for name, ours in (('upfirdn2d', upfirdn2d), ('bias_act', bias_act), ('conv2d_resample', conv2d_resample), ('fma', fma)):
            if hasattr(mod, name):
                monkeypatch.setattr(mod, name, ours)
    return gen


def _load(net, g, prefix, dev):
    sd = {k[len(prefix):]: torch.as_tensor(g[k]) for k in g.files if k.startswith(prefix)}
    net.load_state_dict(sd, strict=True)
    return net.to(dev)


def test_reference_discriminator_on_the_swapped_ops(monkeypatch, golden_cm):
    gen = _reference_cm(monkeypatch)
    from afcm_b200 import _lib
    g = golden_cm
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    D = _load(gen.CoModDiscriminator(**D_CFG).eval(), g, 'D.P.', dev)
    img, c = torch.as_tensor(g['D.img'], device=dev), torch.as_tensor(g['D.c'], device=dev)
    n0 = _lib.launch_count()
    with torch.no_grad():
        y = D(img, c)
    assert _lib.launch_count() - n0 > 20
    assert rel_err(y.cpu().numpy(), g['D.y']) < 1e-4
    # first-order gradients through the native backward kernels (the generator's non-saturating loss term)
    for p in D.parameters():
        p.requires_grad_(True)
    img_g = img.clone().requires_grad_(True)
    torch.nn.functional.softplus(-D(img_g, c)).mean().backward()
    assert rel_err(img_g.grad.cpu().numpy(), g['D.dimg']) < 2e-4
    worst = 0.0
    for k, p in D.named_parameters():
        ref = g['D.G.' + k]
        if np.abs(ref).max() > 0:
            worst = max(worst, rel_err(p.grad.cpu().numpy(), ref))
    print('discriminator parameter gradients: worst rel err %.2e' % worst)
    assert worst < 5e-4


def test_reference_comod_generator_on_the_swapped_ops(monkeypatch, golden_cm):
    gen = _reference_cm(monkeypatch)
    g = golden_cm
    dev = torch.device('cuda:0')
    torch.manual_seed(2)
    G = _load(gen.CoModGenerator(**G_CFG).eval(), g, 'G.P.', dev)
    with torch.no_grad():
        y = G(torch.as_tensor(g['G.z'], device=dev), torch.as_tensor(g['G.c'], device=dev), torch.as_tensor(g['G.x'], device=dev),
              noise_mode='const')
    assert rel_err(y.cpu().numpy(), g['G.y']) < 1e-4


@pytest.mark.parametrize('case', [dict(up=1, down=1, k=3, pad=1), dict(up=1, down=2, k=3, pad=1), dict(up=2, down=1, k=3, pad=1),
                                  dict(up=1, down=2, k=1, pad=0), dict(up=2, down=1, k=1, pad=0), dict(up=1, down=1, k=3, pad=[1, 0, 2, 1]),
                                  dict(up=2, down=2, k=3, pad=1), dict(up=1, down=1, k=3, pad=1, groups=2)])
def test_conv2d_resample_matches_torch(case):
    """The operator itself against the same case analysis evaluated with torch's convolutions (float64 on the CPU)."""
    from afcm_b200.torch_utils.ops import conv2d_resample, upfirdn2d
    dev = torch.device('cuda:0')
    gen = torch.Generator().manual_seed(9)
    groups = case.get('groups', 1)
    x = torch.randn(2, 6, 12, 10, generator=gen)
    w = torch.randn(8, 6 // groups, case['k'], case['k'], generator=gen)
    f = upfirdn2d.setup_filter([1, 3, 3, 1])
    y = conv2d_resample.conv2d_resample(x.to(dev), w.to(dev), f.to(dev), up=case['up'], down=case['down'], padding=case['pad'], groups=groups)
    ref = _resample_ref(x.double(), w.double(), f.double(), case['up'], case['down'], case['pad'], groups)
    assert y.shape == ref.shape
    assert rel_err(y.cpu().numpy(), ref.numpy()) < 1e-5


def _upfirdn_ref(x, f, up=1, down=1, padding=(0, 0, 0, 0), gain=1.0):
    """upfirdn2d by definition (OPS/upfirdn2d.py:167-211) with torch ops: zero insertion, padding / cropping, true convolution."""
    import torch.nn.functional as F
    N, C, H, W = x.shape
    px0, px1, py0, py1 = padding
    x = x.reshape(N, C, H, 1, W, 1)
    x = F.pad(x, [0, up - 1, 0, 0, 0, up - 1])
    x = x.reshape(N, C, H * up, W * up)
    x = F.pad(x, [max(px0, 0), max(px1, 0), max(py0, 0), max(py1, 0)])
    x = x[:, :, max(-py0, 0):x.shape[2] - max(-py1, 0), max(-px0, 0):x.shape[3] - max(-px1, 0)]
    if f is not None:
        f2 = (torch.outer(f, f) if f.ndim == 1 else f) * gain
        f2 = f2.flip([0, 1])[None, None].repeat(C, 1, 1, 1)
        x = F.conv2d(x, f2, groups=C)
    return x[:, :, ::down, ::down]


def _resample_ref(x, w, f, up, down, padding, groups):
    """conv2d_resample by its generic definition (CM conv2d_resample.py:149-153): up-sample + filter, convolve, filter + down-sample."""
    import torch.nn.functional as F
    fw = f.shape[-1]
    p = [padding] * 4 if isinstance(padding, int) else list(padding)
    px0, px1, py0, py1 = p
    if up > 1:
        px0 += (fw + up - 1) // 2; px1 += (fw - up) // 2; py0 += (fw + up - 1) // 2; py1 += (fw - up) // 2
    if down > 1:
        px0 += (fw - down + 1) // 2; px1 += (fw - down) // 2; py0 += (fw - down + 1) // 2; py1 += (fw - down) // 2
    x = _upfirdn_ref(x, f if up > 1 else None, up=up, padding=(px0, px1, py0, py1), gain=up ** 2)
    x = F.conv2d(x, w, groups=groups)
    if down > 1:
        x = _upfirdn_ref(x, f, down=down)
    return x


@pytest.mark.parametrize('k,pad,stride', [(3, 1, 1), (3, 2, 1), (1, 0, 1), (3, 1, 2), (3, 0, 1)])
def test_conv2d_double_backward_matches_torch(k, pad, stride):
    """Second-order gradients of conv2d_gradfix.conv2d (the R1 penalty differentiates the data gradient of every discriminator
    convolution): gradient of |d(sum y r)/dx|^2 with respect to x-independent inputs w and r-weighted y, against torch's own
    double backward in float64."""
    import torch.nn.functional as F
    from afcm_b200.torch_utils.ops import conv2d_gradfix
    dev = torch.device('cuda:0')
    gen = torch.Generator().manual_seed(11)
    x0 = torch.randn(2, 5, 10, 12, generator=gen)
    w0 = torch.randn(7, 5, k, k, generator=gen) * 0.3
    r = torch.randn(2, 7, (10 + 2 * pad - k) // stride + 1, (12 + 2 * pad - k) // stride + 1, generator=gen)

    def penalty(conv, x, w, rr):
        y = conv(x, w)
        gx, = torch.autograd.grad((y * rr).sum() + (y ** 2).sum() * 0.5, x, create_graph=True)
        return gx.square().sum()

    x = x0.double().requires_grad_(True); w = w0.double().requires_grad_(True)
    penalty(lambda a, b: F.conv2d(a, b, stride=stride, padding=pad), x, w, r.double()).backward()
    ref_dx, ref_dw = x.grad.clone(), w.grad.clone()
    xg = x0.to(dev).requires_grad_(True); wg = w0.to(dev).requires_grad_(True)
    penalty(lambda a, b: conv2d_gradfix.conv2d(a, b, stride=stride, padding=pad), xg, wg, r.to(dev)).backward()
    assert rel_err(xg.grad.cpu().numpy(), ref_dx.numpy()) < 1e-4
    assert rel_err(wg.grad.cpu().numpy(), ref_dw.numpy()) < 1e-4


def test_reference_discriminator_r1_penalty(monkeypatch, golden_cm):
    """The R1 regularisation step of the reference (models/comodgan_model.py:128-161) on the swapped operators: the input
    gradient is taken with create_graph=True and its squared norm is differentiated again -- through the native convolution
    (differentiable backward, conv2d_gradfix._ConvFn / _WgradFn), bias_act (second-order kernel) and upfirdn2d."""
    gen = _reference_cm(monkeypatch)
    g = golden_cm
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    D = _load(gen.CoModDiscriminator(**D_CFG).eval(), g, 'D.P.', dev)
    for p in D.parameters():
        p.requires_grad_(True)
    img, c = torch.as_tensor(g['D.img'], device=dev).requires_grad_(True), torch.as_tensor(g['D.c'], device=dev)
    r1_grads = torch.autograd.grad(outputs=[D(img, c).sum()], inputs=[img], create_graph=True, only_inputs=True)[0]
    pen = r1_grads.square().sum([1, 2, 3])
    assert rel_err(pen.detach().cpu().numpy(), g['D.r1_pen']) < 1e-4
    (pen * (10.0 / 2)).mean().backward()
    worst = 0.0
    for k, p in D.named_parameters():
        key = 'D.R1.' + k
        if key in g.files and np.abs(g[key]).max() > 0:
            assert p.grad is not None, k
            worst = max(worst, rel_err(p.grad.cpu().numpy(), g[key]))
    print('R1 penalty parameter gradients: worst rel err %.2e' % worst)
    assert worst < 1e-3
