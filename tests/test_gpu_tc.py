"""tcgen05 / TMEM implicit-GEMM convolution: parity against the exact-fp32 kernel and the reference
golden vectors within the stated 16-bit tolerance (operands rounded to fp16/bf16, fp32 accumulation)."""
import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu

# operands carry 11 (fp16) / 8 (bf16) significant bits; K up to 4608 products accumulate in fp32
TC_TOL = {torch.float16: 2e-3, torch.bfloat16: 1.5e-2}


@pytest.mark.parametrize('dtype', [torch.float16, torch.bfloat16])
@pytest.mark.parametrize('shape', [
    # N, Ci, Co, H, W, pad
    (1, 64, 64, 14, 14, 2),        # single k-block per tap, single tile
    (2, 128, 96, 36, 36, 2),       # AFCM 36x36 plane, two channel blocks
    (1, 4, 64, 52, 52, 2),         # enc0-like: Ci << 64 (TMA zero-fills the channel tail)
    (2, 91, 181, 30, 22, 2),       # odd channel counts on both sides, rectangular plane
    (1, 362, 512, 38, 38, 2),      # Co tiled as 2 x 256, Ci tail
    (2, 512, 362, 20, 20, 2),      # Co tiled as 2 x 192
    (2, 96, 80, 36, 36, 1),        # pad 1 (e_16x16): garbage columns are skipped on store
    (3, 48, 45, 52, 52, 2),        # tiny-generator shapes
])
def test_conv2d_tc_matches_fp32(shape, dtype):
    from afcm_b200.torch_utils.ops import conv2d_gradfix
    N, Ci, Co, H, W, pad = shape
    dev = torch.device('cuda:0')
    g = torch.Generator(device='cpu').manual_seed(17)
    x = torch.randn(N, Ci, H, W, generator=g).to(dev)
    w = torch.randn(Co, Ci, 3, 3, generator=g).to(dev)
    icoef = (torch.rand(N, Ci, generator=g) + 0.5).to(dev)
    ocoef = (torch.rand(N, Co, generator=g) + 0.5).to(dev)
    scale = 1.0 / np.sqrt(Ci * 9)
    conv2d_gradfix.set_conv_impl('f32', dtype)
    ref = conv2d_gradfix.conv2d_native(x, w, pad, icoef=icoef, ocoef=ocoef, pre_scale=scale, impl='f32')
    got = conv2d_gradfix.conv2d_native(x, w, pad, icoef=icoef, ocoef=ocoef, pre_scale=scale, impl='tc')
    torch.cuda.synchronize()
    assert got.shape == ref.shape
    assert torch.isfinite(got).all()
    assert rel_err(got.cpu().numpy(), ref.cpu().numpy()) < TC_TOL[dtype]
    conv2d_gradfix.set_conv_impl('f32', torch.float16)


def test_conv2d_tc_exact_on_representable_inputs():
    """With inputs exactly representable in fp16 and small integer products the tensor-core result must
    be bit-identical to the fp32 kernel: proves descriptor / layout correctness, not just 'close'."""
    from afcm_b200.torch_utils.ops import conv2d_gradfix
    dev = torch.device('cuda:0')
    g = torch.Generator(device='cpu').manual_seed(3)
    N, Ci, Co, H, W = 2, 192, 160, 36, 36
    x = torch.randint(-4, 5, (N, Ci, H, W), generator=g).float().to(dev)
    w = torch.randint(-2, 3, (Co, Ci, 3, 3), generator=g).float().to(dev)
    ref = conv2d_gradfix.conv2d_native(x, w, 2, impl='f32')
    got = conv2d_gradfix.conv2d_native(x, w, 2, impl='tc')
    assert torch.equal(ref, got)


def test_modulated_conv2d_golden_tc(golden_ops):
    from afcm_b200.networks_stylegan3 import modulated_conv2d
    dev = torch.device('cuda:0')
    g = golden_ops
    for name in ('demod3', 'demod3b'):
        t = 'modconv.' + name
        demod, pad, ig = int(g[t + '.cfg'][0]), int(g[t + '.cfg'][1]), float(g[t + '.cfg'][2])
        y = modulated_conv2d(torch.as_tensor(g[t + '.x'], device=dev), torch.as_tensor(g[t + '.w'], device=dev),
                             torch.as_tensor(g[t + '.s'], device=dev), bool(demod), pad, torch.tensor(ig, device=dev), impl='tc')
        assert rel_err(y.cpu().numpy(), g[t + '.y']) < 2e-3, name


def test_full_generator_tc(golden_full):
    """Whole generator on the tensor-core path vs the reference fp32 output: stated tolerance
    PSNR >= 50 dB (peak = max|y_ref|) and max-abs error <= 1% of max|y_ref| with fp16 operands."""
    from afcm_b200.networks_stylegan3 import afcm_generator
    from afcm_b200.torch_utils.ops import conv2d_gradfix
    dev = torch.device('cuda:0')
    g = golden_full
    G = afcm_generator(seed=0, device=dev)
    x = (torch.as_tensor(g['x_u8']).float() * (2.0 / 255.0) - 1.0).clamp(-1, 1).to(dev)
    conv2d_gradfix.set_conv_impl('tc', torch.float16)
    try:
        with torch.no_grad():
            y = G(torch.as_tensor(g['z'], device=dev), torch.as_tensor(g['c'], device=dev), x, noise_mode='const')
    finally:
        conv2d_gradfix.set_conv_impl('f32', torch.float16)
    ref = g['y']
    err = y.cpu().numpy() - ref
    peak = np.abs(ref).max()
    psnr = 10 * np.log10(peak ** 2 / np.mean(err ** 2))
    print(f'tc fp16: psnr {psnr:.1f} dB, max-abs/peak {np.abs(err).max() / peak:.2e}')
    assert psnr >= 50.0
    assert np.abs(err).max() / peak <= 1e-2


@pytest.mark.parametrize('shape', [(2, 128, 96, 36, 36, 2), (1, 4, 64, 52, 52, 2), (2, 91, 181, 30, 22, 2),
                                   (2, 96, 80, 36, 36, 1), (1, 8, 16, 21, 21, 2)])
def test_conv2d_tc_fp16_storage(shape):
    """fp16 activations in, fp16 result out (the storage format of the fast inference path): the same GEMM,
    so the result equals the fp32-in / fp32-out tensor-core result of the fp16-rounded input, rounded once."""
    from afcm_b200.torch_utils.ops import conv2d_gradfix
    N, Ci, Co, H, W, pad = shape
    dev = torch.device('cuda:0')
    g = torch.Generator(device='cpu').manual_seed(23)
    x = torch.randn(N, Ci, H, W, generator=g).to(dev).half()
    w = torch.randn(Co, Ci, 3, 3, generator=g).to(dev)
    icoef = (torch.rand(N, Ci, generator=g) + 0.5).to(dev)
    ocoef = (torch.rand(N, Co, generator=g) + 0.5).to(dev)
    scale = 1.0 / np.sqrt(Ci * 9)
    conv2d_gradfix.set_conv_impl('f32', torch.float16)
    ref = conv2d_gradfix.conv2d_native(x.float(), w, pad, icoef=icoef, ocoef=ocoef, pre_scale=scale, impl='tc')
    got = conv2d_gradfix.conv2d_native(x, w, pad, icoef=icoef, ocoef=ocoef, pre_scale=scale, impl='tc',
                                       out_dtype=torch.float16)
    assert got.dtype == torch.float16 and got.shape == ref.shape
    assert torch.equal(got, ref.half())          # odd W (21) exercises the scalar pack kernel
    exact = conv2d_gradfix.conv2d_native(x.float(), w, pad, icoef=icoef, ocoef=ocoef, pre_scale=scale, impl='f32')
    assert rel_err(got.float().cpu().numpy(), exact.cpu().numpy()) < 3e-3


@pytest.mark.parametrize('mode', [0, 1, 2, 4])
def test_conv2d_tc_rowreuse_modes_exact(mode):
    """All pipelines of the tcgen05 kernel -- one A tile per tap (0); one A tile per kernel row re-read at a 128-byte
    descriptor offset per kx tap, with the weights resident in shared memory where they fit (1), streamed (2), or streamed through a ring of their
    own (4) -- are
    bit-identical to the fp32 kernel on representable inputs."""
    from afcm_b200 import _lib
    from afcm_b200.torch_utils.ops import conv2d_gradfix
    dev = torch.device('cuda:0')
    g = torch.Generator(device='cpu').manual_seed(5)
    _lib.lib().afcm_conv_tc_set_rowreuse(mode)
    try:
        for (N, Ci, Co, H, W) in [(2, 64, 64, 36, 36), (1, 192, 96, 30, 22), (2, 91, 128, 52, 52), (1, 200, 181, 20, 38), (1, 128, 256, 36, 36),
                                  (1, 96, 300, 20, 20)]:
            x = torch.randint(-4, 5, (N, Ci, H, W), generator=g).float().to(dev)
            w = torch.randint(-2, 3, (Co, Ci, 3, 3), generator=g).float().to(dev)
            ref = conv2d_gradfix.conv2d_native(x, w, 2, impl='f32')
            got = conv2d_gradfix.conv2d_native(x, w, 2, impl='tc')
            assert torch.equal(ref, got), (mode, N, Ci, Co, H, W)
            ref1 = conv2d_gradfix.conv2d_native(x, w, 1, impl='f32')
            got1 = conv2d_gradfix.conv2d_native(x, w, 1, impl='tc')
            assert torch.equal(ref1, got1), (mode, 'pad1', N, Ci, Co, H, W)
    finally:
        _lib.lib().afcm_conv_tc_set_rowreuse(-1)


@pytest.mark.parametrize('shape', [(1, 128, 128, 276, 276), (1, 181, 181, 276, 276), (1, 128, 96, 276, 276), (1, 64, 64, 276, 276)])
def test_conv2d_tc_direct_nchw_is_repeatable(shape):
    """Regression test of a shared-memory race in the direct-NCHW convolution: the producer warps released a raw TMA slot
    when their loads from it had been issued, not when they had returned; with the shallow operand rings of the 96..192
    channel layers the refill then landed under the last loads (a few dozen wrong pixels per launch, timing dependent).
    Many tiles per CTA, several launches, each bit-identical to the packed path -- with one and with two MMA issuers."""
    from afcm_b200 import _lib
    from afcm_b200.torch_utils.ops import conv2d_gradfix
    N, Ci, Co, H, W = shape
    dev = torch.device('cuda:0')
    g = torch.Generator(device='cpu').manual_seed(5)
    x = torch.randn(N, Ci, H, W, generator=g).to(dev).half()
    w = torch.randn(Co, Ci, 3, 3, generator=g).to(dev)
    bias = torch.randn(Co, generator=g).to(dev)
    conv2d_gradfix.set_conv_impl('f32', torch.float16)
    run = lambda: conv2d_gradfix.conv2d_native(x, w, 2, pre_scale=1.0 / np.sqrt(Ci * 9), impl='tc', out_dtype=torch.float16, bias=bias)
    try:
        conv2d_gradfix.hybrid_pack = False                  # every shape through the direct kernel
        conv2d_gradfix.direct_nchw = False
        _lib.lib().afcm_conv_tc_set_issuers(1)
        ref = run()
        for issuers in (1, 2):
            _lib.lib().afcm_conv_tc_set_issuers(issuers)
            for direct in (False, True):
                conv2d_gradfix.direct_nchw = direct
                for rep in range(6):
                    got = run()
                    assert torch.equal(got, ref), (issuers, direct, rep, int((got != ref).sum()))
    finally:
        conv2d_gradfix.direct_nchw = True
        conv2d_gradfix.hybrid_pack = False
        _lib.lib().afcm_conv_tc_set_issuers(1)


def test_conv2d_tc_hybrid_packs_pitched_planes():
    """Few input channels on large planes: conv2d_native takes pack + GEMM instead of the direct kernel (conv2d_gradfix._prefers_pack),
    and the pack kernel reads the planes at the W + 2 pitch filtered_lrelu_tc wrote them at -- no dense copy in between."""
    from afcm_b200.torch_utils.ops import conv2d_gradfix
    N, Ci, Co, H, W = 2, 64, 91, 276, 276
    dev = torch.device('cuda:0')
    g = torch.Generator(device='cpu').manual_seed(9)
    x = torch.randn(N, Ci, H, W, generator=g).to(dev).half()
    w = torch.randn(Co, Ci, 3, 3, generator=g).to(dev)
    icoef = (torch.rand(N, Ci, generator=g) + 0.5).to(dev)
    bias = torch.randn(Co, generator=g).to(dev)
    conv2d_gradfix.set_conv_impl('f32', torch.float16)
    run = lambda t: conv2d_gradfix.conv2d_native(t, w, 2, icoef=icoef, pre_scale=1.0 / np.sqrt(Ci * 9), impl='tc', out_dtype=torch.float16, bias=bias)
    default = conv2d_gradfix.hybrid_pack
    conv2d_gradfix.hybrid_pack = True
    assert conv2d_gradfix._prefers_pack(Ci, Co, H, W) and not conv2d_gradfix._prefers_pack(128, Co, H, W)
    ref = run(x)
    xq = torch.full((N, Ci, H, W + 2), float('nan'), device=dev, dtype=torch.float16)      # the pad columns must not be read
    xq[..., :W] = x
    n0 = _lib_launches()
    got = run(xq[..., :W])
    assert _lib_launches() - n0 == 2                        # pack + GEMM, no copy kernel of ours in between
    assert torch.equal(got, ref)
    try:
        conv2d_gradfix.hybrid_pack = False
        xz = torch.zeros_like(xq); xz[..., :W] = x
        assert torch.equal(run(xz[..., :W]), ref)           # and the direct kernel agrees bit for bit
    finally:
        conv2d_gradfix.hybrid_pack = default


@pytest.mark.parametrize('shape', [(2, 64, 64, 36, 36), (1, 4, 64, 52, 52), (2, 91, 181, 30, 22), (1, 181, 256, 38, 36), (2, 512, 512, 36, 36),
                                   (1, 362, 512, 20, 84), (1, 128, 96, 276, 276), (1, 96, 300, 20, 20), (3, 8, 16, 8, 130)])
@pytest.mark.parametrize('mod', [False, True])
def test_conv2d_tc_direct_nchw_is_bit_identical_to_pack(shape, mod):
    """afcm_conv2d_tc_nchw (producer warps build the A tiles from the fp16 NCHW planes, modulation applied on the
    way) returns exactly what afcm_conv_tc_pack + afcm_conv2d_tc return -- every pipeline of the kernel (resident weights,
    combined stages, separate A / B rings), ragged channel counts, planes from 8 to 276 pixels, several rows per tile."""
    from afcm_b200.torch_utils.ops import conv2d_gradfix
    N, Ci, Co, H, W = shape
    dev = torch.device('cuda:0')
    g = torch.Generator(device='cpu').manual_seed(31)
    x = torch.randn(N, Ci, H, W, generator=g).to(dev).half()
    w = torch.randn(Co, Ci, 3, 3, generator=g).to(dev)
    icoef = (torch.rand(N, Ci, generator=g) + 0.5).to(dev) if mod else None
    ocoef = (torch.rand(N, Co, generator=g) + 0.5).to(dev) if mod else None
    bias = torch.randn(Co, generator=g).to(dev)
    scale = 1.0 / np.sqrt(Ci * 9)
    conv2d_gradfix.set_conv_impl('f32', torch.float16)
    run = lambda: conv2d_gradfix.conv2d_native(x, w, 2, icoef=icoef, ocoef=ocoef, pre_scale=scale, impl='tc', out_dtype=torch.float16, bias=bias)
    try:
        conv2d_gradfix.hybrid_pack = False
        conv2d_gradfix.direct_nchw = False
        ref = run()
        conv2d_gradfix.direct_nchw = True
        n0 = _lib_launches()
        got = run()
        assert _lib_launches() - n0 == (1 if (H * W) % 8 == 0 else 2)          # one kernel: no pack launch (H W % 8: TMA stride rule)
        # the same planes stored at the row pitch W + 2 with zero pad columns (what filtered_lrelu_tc writes for the convolution)
        xq = torch.zeros(N, Ci, H, W + 2, device=dev, dtype=torch.float16)
        xq[..., :W] = x
        x = xq[..., :W]
        n0 = _lib_launches()
        got2 = run()
        assert _lib_launches() - n0 == (1 if (H * (W + 2)) % 8 == 0 else 2)
    finally:
        conv2d_gradfix.direct_nchw = True
        conv2d_gradfix.hybrid_pack = False
    assert got.shape == ref.shape and torch.equal(got, ref)
    assert torch.equal(got2, ref)


def _lib_launches():
    from afcm_b200 import _lib
    return _lib.launch_count()
