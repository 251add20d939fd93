"""GPU parity of the tensor-core filtered_lrelu kernels with sign tensor (the training-step variants: 'tcs' =
afcm_filtered_lrelu_tcs, shared-memory tiled, bf16 backward; 'tc' = afcm_filtered_lrelu_tc_signs, the register-chained kernel
with the sign tensor, fp16 backward scaled by max|dy|):
forward against the reference golden vectors and against the exact fp32 kernel, sign tensors against the exact kernel's,
backward (sign read, bf16 operands) against the reference's autograd gradients.  Stated tolerances: forward 2e-3 of
max|y| (fp16 operands); sign codes equal except where the up-sampled value is within fp16 rounding distance of zero or
of the clamp (< 0.5 % of the codes on random data); backward: relative L2 error < 3e-2 -- a flipped sign changes that
sample's gradient by a factor 1/slope, so the max-norm error is set by the few flipped samples (bounded at 0.15 of
max|dx|), while bf16 operand rounding alone gives ~5e-3."""
import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu

FWD_TOL, BWD_L2, BWD_MAX = 2e-3, 3e-2, 0.15


def l2_err(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


@pytest.fixture(autouse=True)
def _impl():
    from afcm_b200.torch_utils.ops import filtered_lrelu
    yield
    filtered_lrelu.set_train_impl('exact')


def _case(g, name, dev):
    k = 'flrelu.' + name
    cfg = g[k + '.cfg']
    up, dn = int(cfg[0]), int(cfg[1])
    pad = [int(v) for v in cfg[2:6]]
    gain, slope = float(cfg[6]), float(cfg[7])
    clamp = None if cfg[8] < 0 else float(cfg[8])
    fu = torch.as_tensor(g[k + '.fu'], device=dev) if g[k + '.fu'].size else None
    fd = torch.as_tensor(g[k + '.fd'], device=dev) if g[k + '.fd'].size else None
    return k, up, dn, pad, gain, slope, clamp, fu, fd


@pytest.mark.parametrize('impl', ['tcs', 'tc'])
@pytest.mark.parametrize('name', ['u2d2', 'u2d4', 'u4d2', 'crop', 'clamp', 'rect'])
def test_forward_backward_vs_reference_golden(golden_ops, name, impl):
    from afcm_b200.torch_utils.ops import filtered_lrelu
    dev = torch.device('cuda:0')
    g = golden_ops
    k, up, dn, pad, gain, slope, clamp, fu, fd = _case(g, name, dev)
    filtered_lrelu.set_train_impl(impl)
    x = torch.as_tensor(g[k + '.x'], device=dev).requires_grad_(True)
    b = torch.as_tensor(g[k + '.b'], device=dev).requires_grad_(True)
    y = filtered_lrelu.filtered_lrelu(x, fu=fu, fd=fd, b=b, up=up, down=dn, padding=pad, gain=gain, slope=slope, clamp=clamp)
    assert y.shape == g[k + '.y'].shape
    assert rel_err(y.detach().cpu().numpy(), g[k + '.y']) < FWD_TOL
    (y * torch.as_tensor(g[k + '.r'], device=dev)).sum().backward()
    dx, db = x.grad.cpu().numpy(), b.grad.cpu().numpy()
    print(name, 'fwd', rel_err(y.detach().cpu().numpy(), g[k + '.y']), 'dx l2', l2_err(dx, g[k + '.dx']), 'max', rel_err(dx, g[k + '.dx']),
          'db', rel_err(db, g[k + '.db']))
    assert l2_err(dx, g[k + '.dx']) < BWD_L2 and rel_err(dx, g[k + '.dx']) < BWD_MAX
    assert rel_err(db, g[k + '.db']) < BWD_L2


@pytest.mark.parametrize('impl', ['tcs', 'tc'])
@pytest.mark.parametrize('geo', [(2, 2, 12, 12, [9, 8, 9, 8], 278), (2, 4, 12, 24, [34, 33, 34, 33], 150), (4, 2, 24, 12, [-6, -9, -6, -9], 86),
                                 (2, 2, 12, 12, [-11, -12, -11, -12], 278), (2, 2, 12, 12, [9, 8, 9, 8], 38), (2, 4, 12, 24, [34, 33, 34, 33], 278),
                                 (4, 2, 24, 12, [-6, -9, -6, -9], 150), (2, 4, 12, 24, [34, 33, 34, 33], 54), (4, 2, 24, 12, [-6, -9, -6, -9], 38)])
def test_multi_tile_vs_exact_kernel(geo, impl):
    """Planes of the AFCM layer sizes (several tiles, ragged edges): forward, sign tensor and backward against the exact kernel."""
    from afcm_b200.networks_stylegan3 import design_lowpass_filter
    from afcm_b200.torch_utils.ops import filtered_lrelu
    up, dn, nfu, nfd, pad, H = geo
    dev = torch.device('cuda:0')
    gen = torch.Generator().manual_seed(2)
    fu = design_lowpass_filter(nfu, 40.0, 30.0, 512).to(dev)
    fd = design_lowpass_filter(nfd, 30.0, 30.0, 512).to(dev)
    x0 = (torch.randn(2, 3, H, H, generator=gen) * 2).to(dev)
    b0 = (torch.randn(3, generator=gen) * 0.3).to(dev)
    res = {}
    n_tc = filtered_lrelu.tc_sign_calls
    for which in ('exact', impl):
        filtered_lrelu.set_train_impl(which)
        x = x0.clone().requires_grad_(True); b = b0.clone().requires_grad_(True)
        fn = filtered_lrelu._filtered_lrelu_cuda(up=up, down=dn, padding=pad, gain=np.sqrt(2), slope=0.2, clamp=4.0)
        y = fn.apply(x, fu, fd, b, None, 0, 0)
        signs = y.grad_fn.saved_tensors[2].clone()
        r = torch.randn(y.shape, generator=torch.Generator().manual_seed(3)).to(dev)
        (y * r).sum().backward()
        res[which] = (y.detach().cpu().numpy(), signs.cpu().numpy(), x.grad.cpu().numpy(), b.grad.cpu().numpy())
        ye_shape = tuple(y.shape)
    if impl == 'tc':                                             # the forward always, the backward where the dispatch prefers it
        assert filtered_lrelu.tc_sign_calls - n_tc == (2 if (up, dn) == (2, 2) or H < 48 else 1)      # (the backward's output is x-sized)
    ye, se, dxe, dbe = res['exact']
    yt, st, dxt, dbt = res[impl]
    assert rel_err(yt, ye) < FWD_TOL
    assert se.shape == st.shape
    # codes of the active columns only: the sign tensor width is padded to 16 elements and the padding is never written
    sw = ye.shape[3] * dn - (dn - 1) + (nfd - 1)
    ce = np.stack([(se >> sft) & 3 for sft in range(0, 8, 2)], axis=-1).reshape(se.shape[0], se.shape[1], se.shape[2], -1)[..., :sw]
    ct = np.stack([(st >> sft) & 3 for sft in range(0, 8, 2)], axis=-1).reshape(st.shape[0], st.shape[1], st.shape[2], -1)[..., :sw]
    frac = float((ce != ct).mean())
    assert frac < 5e-3, frac
    print(geo, impl, 'fwd', rel_err(yt, ye), 'sign diff', frac, 'dx l2', l2_err(dxt, dxe), 'max', rel_err(dxt, dxe), 'db', rel_err(dbt, dbe))
    assert l2_err(dxt, dxe) < BWD_L2 and rel_err(dxt, dxe) < BWD_MAX
    assert rel_err(dbt, dbe) < BWD_L2


def test_signs_interoperate_with_exact_kernel():
    """Same sign tensor format: tensor-core forward + exact backward == exact forward + exact backward up to the fp16 signs."""
    from afcm_b200.networks_stylegan3 import design_lowpass_filter
    from afcm_b200.torch_utils.ops import filtered_lrelu
    dev = torch.device('cuda:0')
    fu = design_lowpass_filter(12, 40.0, 30.0, 512).to(dev)
    x0 = torch.randn(1, 2, 70, 70, generator=torch.Generator().manual_seed(4)).to(dev)
    outs = []
    for fwd_impl in ('exact', 'tc', 'tcs'):
        filtered_lrelu.set_train_impl(fwd_impl)
        x = x0.clone().requires_grad_(True)
        y = filtered_lrelu.filtered_lrelu(x, fu=fu, fd=fu, b=None, up=2, down=2, padding=[9, 8, 9, 8], clamp=256)
        filtered_lrelu.set_train_impl('exact')               # backward always on the exact kernel
        y.sum().backward()
        outs.append(x.grad.cpu().numpy())
    assert rel_err(outs[1], outs[0]) < 2e-2 and rel_err(outs[2], outs[0]) < 2e-2


@pytest.mark.parametrize('geo', [(2, 2, 12, 12, [9, 8, 9, 8], 86), (2, 4, 12, 24, [34, 33, 34, 33], 38), (4, 2, 24, 12, [-6, -9, -6, -9], 22)])
@pytest.mark.parametrize('scale', [1e-9, 1e-4, 3e4])
def test_backward_gradients_far_from_the_fp16_range(geo, scale):
    """The fp16 backward of the register-chained kernel scales its operands by a power of two derived from max|dy| (afcm_absmax):
    gradients of 1e-9 (below the smallest fp16 subnormal) and of 3e4 (overflowing after the up-sampling gain) come out like the
    exact kernel's."""
    from afcm_b200.networks_stylegan3 import design_lowpass_filter
    from afcm_b200.torch_utils.ops import filtered_lrelu
    up, dn, nfu, nfd, pad, H = geo
    dev = torch.device('cuda:0')
    fu = design_lowpass_filter(nfu, 40.0, 30.0, 512).to(dev)
    fd = design_lowpass_filter(nfd, 30.0, 30.0, 512).to(dev)
    x0 = (torch.randn(2, 4, H, H, generator=torch.Generator().manual_seed(5)) * 2).to(dev)
    out = {}
    for which in ('exact', 'tc'):
        filtered_lrelu.set_train_impl(which)
        x = x0.clone().requires_grad_(True)
        y = filtered_lrelu.filtered_lrelu(x, fu=fu, fd=fd, b=None, up=up, down=dn, padding=pad, gain=np.sqrt(2), slope=0.2, clamp=256)
        r = torch.randn(y.shape, generator=torch.Generator().manual_seed(6)).to(dev) * scale
        (y * r).sum().backward()
        out[which] = x.grad.cpu().numpy()
    assert np.isfinite(out['tc']).all()
    assert l2_err(out['tc'], out['exact']) < BWD_L2, (l2_err(out['tc'], out['exact']), scale)
