"""Size-independent properties of the hot path at the BASELINE geometries (full sizes are too large for the CPU
oracle, so the kernels are checked through identities that must hold BIT-EXACTLY whatever the size):

* positive homogeneity with a power-of-two factor: scaling fp16 / fp32 data by 2 commutes with every rounding of the pipeline as
  long as nothing overflows, becomes subnormal or reaches the clamp, so conv(2 x) == 2 conv(x) bit for bit and, because leaky
  ReLU is positively homogeneous, filtered_lrelu(2 x) == 2 filtered_lrelu(x) (bit for bit except where an fp16 intermediate of
  the kernel is subnormal; to the kernel's tolerance everywhere);
* planes / samples are independent: permuting the batch (or the channels of filtered_lrelu) permutes the result -- which also
  exercises a different assignment of tiles to CTAs, i.e. the result must not depend on the schedule;
* a batch is the concatenation of its slices: the generator on a batch equals the generator on its halves.
Every test runs the benchmarked kernels (tcgen05 convolution reading fp16 planes directly, tensor-core filtered_lrelu)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _away_from_zero(shape, gen):
    """fp16 test data with 0.25 <= |x| < 1.25: no product or sum operand of the pipeline is subnormal in fp16, where a factor of 2
    would not commute with the rounding."""
    mag = torch.rand(shape, generator=gen) + 0.25
    sign = (torch.rand(shape, generator=gen) < 0.5).float() * 2 - 1
    return (mag * sign).half()


def _assert_doubled(y2, y):
    """y2 == 2 y bit for bit wherever y is not tiny (a result below ~1e-3 may come out of a subnormal intermediate); everywhere
    else within one fp16 subnormal step."""
    big = y.abs() > 1e-3
    assert torch.equal(y2[big], (y * 2)[big]), int((y2[big] != (y * 2)[big]).sum())
    assert float((y2.float() - 2 * y.float())[~big].abs().max() if (~big).any() else 0.0) <= 2.0 ** -22


def _layer(name):
    from afcm_b200.networks_stylegan3 import afcm_generator
    G = afcm_generator(seed=0, device=torch.device('cuda:0'))
    return G, getattr(G.synthesis, name)


@pytest.mark.parametrize('shape', [(6, 64, 64, 276), (4, 128, 181, 276), (6, 362, 512, 148), (8, 512, 512, 84), (16, 512, 512, 36)])
@pytest.mark.parametrize('pitched', [False, True])
def test_convolution_homogeneity_and_batch_permutation(shape, pitched):
    from afcm_b200.torch_utils.ops import conv2d_gradfix
    N, Ci, Co, H = shape
    dev = torch.device('cuda:0')
    g = torch.Generator().manual_seed(H + Ci)
    x = _away_from_zero((N, Ci, H, H), g).to(dev)
    if pitched:                                             # the hand-over format filtered_lrelu_tc writes (row pitch W + 2, zero pad columns)
        xq = torch.zeros(N, Ci, H, H + 2, device=dev, dtype=torch.float16)
        xq[..., :H] = x
        x = xq[..., :H]
    w = torch.randn(Co, Ci, 3, 3, generator=g).to(dev)
    icoef = (torch.rand(N, Ci, generator=g) + 0.5).to(dev)
    ocoef = (torch.rand(N, Co, generator=g) + 0.5).to(dev)
    bias = torch.randn(Co, generator=g).to(dev)
    conv2d_gradfix.set_conv_impl('f32', torch.float16)
    run = lambda t, ic, b: conv2d_gradfix.conv2d_native(t, w, 2, icoef=ic, ocoef=ocoef, pre_scale=1.0 / np.sqrt(Ci * 9), impl='tc',
                                                        out_dtype=torch.float16, bias=b)
    y = run(x, icoef, None)
    assert torch.isfinite(y).all() and float(y.abs().max()) < 3e4
    two_x = x * 2 if not pitched else (xq * 2)[..., :H]
    _assert_doubled(run(two_x, icoef, None), y)                              # activations scaled
    _assert_doubled(run(x, icoef * 2, None), y)                              # modulation scaled
    perm = torch.randperm(N, generator=g).to(dev)
    xp = x[perm] if not pitched else xq[perm][..., :H]
    yb = run(x, icoef, bias)
    ocoef_saved = ocoef
    ocoef = ocoef_saved[perm]
    assert torch.equal(run(xp, icoef[perm], bias), yb[perm])                 # samples are independent, whatever CTA computes them


@pytest.mark.parametrize('case', [(2, 2, 4, 64, 278), (2, 4, 3, 181, 278), (2, 2, 4, 256, 150), (2, 4, 2, 512, 150), (4, 2, 4, 362, 86),
                                  (4, 2, 3, 128, 150), (2, 2, 8, 512, 38)])
def test_filtered_lrelu_homogeneity_and_plane_permutation(case):
    from afcm_b200.networks_stylegan3 import design_lowpass_filter
    from afcm_b200.torch_utils.ops.filtered_lrelu import filtered_lrelu_tc
    up, down, N, C, H = case
    dev = torch.device('cuda:0')
    taps = {2: 12, 4: 24}
    fu = design_lowpass_filter(taps[up], 64.0, 30.0, 512).to(dev)
    fd = design_lowpass_filter(taps[down], 64.0 / down * 2, 30.0, 512).to(dev)
    pad = {(2, 2): [9, 8, 9, 8], (2, 4): [34, 33, 34, 33], (4, 2): [-6, -9, -6, -9]}[(up, down)]
    g = torch.Generator().manual_seed(C + H)
    x = _away_from_zero((N, C, H, H), g).to(dev)
    run = lambda t: filtered_lrelu_tc(t, fu, fd, None, up=up, down=down, padding=pad, gain=float(np.sqrt(2)), slope=0.2, clamp=256.0,
                                      out_dtype=torch.float16)
    y = run(x)
    assert y is not None and torch.isfinite(y).all() and float(y.abs().max()) < 64      # far from the clamp (256) and from fp16 overflow
    y2 = run(x * 2)
    # the kernel computes in units of the clamp and accumulates the first three passes in fp16: an intermediate sample below
    # 6e-5 * clamp = 0.016 is subnormal there and rounds differently after doubling (~1 % of the samples of this data), so the
    # identity holds to the kernel's stated tolerance, bit-exactly for the rest
    frac = float((y2 != y * 2).float().mean())
    err = float((y2.float() - 2 * y.float()).abs().max()) / (2 * float(y.abs().max()))
    print(f'filtered_lrelu up {up} down {down} {H} px: {frac:.2e} of the samples differ after doubling, max rel err {err:.1e}')
    assert frac < 0.05 and err <= 2e-3
    perm = torch.randperm(N * C, generator=g).to(dev)
    xp = x.reshape(N * C, H, H)[perm].reshape(N, C, H, H)
    yp = run(xp)
    assert torch.equal(yp.reshape(N * C, *y.shape[2:]), y.reshape(N * C, *y.shape[2:])[perm])


def test_generator_batch_is_the_concatenation_of_its_slices():
    """Slices are independent in the encoder (bit-exact under splitting and permuting the batch).  The synthesis layers are NOT
    exactly independent in the reference either: modulated_conv2d pre-normalises the styles by their mean square over the WHOLE
    [batch, channel] tensor (NET:43); demodulation cancels that factor mathematically, not in floating point, so their outputs
    agree to rounding (measured: one or two fp16 steps per layer)."""
    from afcm_b200 import inference
    from afcm_b200.networks_stylegan3 import afcm_generator
    dev = torch.device('cuda:0')
    G = afcm_generator(seed=0, device=dev)
    S = G.synthesis
    taps = {}
    for i in range(S.num_layers):
        getattr(S, f'encoder_{i}').register_forward_hook(lambda m, a, o, i=i: taps.__setitem__(f'enc{i}', o.detach().clone()))
    g = torch.Generator().manual_seed(21)
    B = 8
    z = torch.randn(B, 512, generator=g).to(dev); c = torch.rand(B, 1, generator=g).to(dev)
    x = (torch.rand(B, 4, 256, 256, generator=g) * 2 - 1).to(dev)
    inference.set_precision('fast')
    try:
        with torch.no_grad():
            y = G(z, c, x, noise_mode='const'); full = dict(taps)
            ya = G(z[:3], c[:3], x[:3], noise_mode='const'); ta = dict(taps)
            yb = G(z[3:], c[3:], x[3:], noise_mode='const'); tb = dict(taps)
            perm = torch.randperm(B, generator=g).to(dev)
            yp = G(z[perm], c[perm], x[perm], noise_mode='const'); tp = dict(taps)
    finally:
        inference.set_precision('fp32')
    for k, v in full.items():
        assert torch.equal(torch.cat([ta[k], tb[k]]), v), k
        assert torch.equal(tp[k], v[perm]), k
    peak = float(y.abs().max())
    assert float((torch.cat([ya, yb]) - y).abs().max()) <= 5e-3 * peak
    assert float((yp - y[perm]).abs().max()) <= 5e-3 * peak
