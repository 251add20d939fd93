"""GPU parity of the tensor-core filtered_lrelu (afcm_filtered_lrelu_tc) against the CPU oracle and against the
exact-fp32 kernel.  Tolerance of this path (fp16 operands, fp32 accumulation): max|err| <= 2e-3 * max|y|."""
import numpy as np
import pytest
import scipy.signal

pytestmark = pytest.mark.gpu
TOL = 2e-3


def _filters(up, down):
    fu = scipy.signal.firwin(6 * up, 0.4, width=0.3, fs=2).astype(np.float32)
    fd = scipy.signal.firwin(6 * down, 0.25, width=0.2, fs=2).astype(np.float32)
    return fu, fd


CASES = [  # up, down, padding, N, C, H, W
    (2, 2, [9, 8, 9, 8], 2, 3, 22, 26),
    (2, 2, [9, 8, 9, 8], 1, 5, 38, 38),
    (2, 4, [34, 33, 34, 33], 2, 3, 38, 42),
    (2, 4, [34, 33, 34, 33], 1, 2, 54, 54),
    (4, 2, [-6, -9, -6, -9], 2, 3, 22, 26),
    (4, 2, [-6, -9, -6, -9], 1, 4, 38, 38),
    (2, 2, [-11, -12, -11, -12], 1, 3, 38, 36),
    (2, 2, [9, 8, 7, 10], 2, 2, 21, 20),
    (4, 2, [3, 2, 1, 4], 1, 3, 9, 12),
    (2, 2, [9, 8, 9, 8], 1, 2, 150, 150),
    (2, 2, [8, 9, 10, 7], 1, 2, 70, 84),          # odd phase shift on x
    (2, 4, [33, 34, 35, 32], 1, 2, 86, 86),
    (4, 2, [-5, -10, -7, -8], 1, 2, 54, 54),
]
CLAMPS = [2.0, None, 4096.0]     # sat() form, no clamp, explicit min/max form


@pytest.mark.parametrize('clamp', CLAMPS)
@pytest.mark.parametrize('up,down,pad,N,C,H,W', CASES)
def test_tc_matches_oracle(up, down, pad, N, C, H, W, clamp):
    import torch
    from afcm_b200.torch_utils.ops.filtered_lrelu import filtered_lrelu_tc
    from oracle import afcm_oracle as orc
    rng = np.random.RandomState(H * 7 + W)
    fu, fd = _filters(up, down)
    x = (rng.randn(N, C, H, W) * 2).astype(np.float32)
    b = rng.randn(C).astype(np.float32)
    ref = orc.filtered_lrelu(x, fu, fd, b, up=up, down=down, padding=pad, gain=np.sqrt(2), slope=0.2, clamp=clamp)
    ref = ref[0] if isinstance(ref, tuple) else ref
    dev = torch.device('cuda')
    y = filtered_lrelu_tc(torch.from_numpy(x).to(dev), torch.from_numpy(fu).to(dev), torch.from_numpy(fd).to(dev),
                          torch.from_numpy(b).to(dev), up=up, down=down, padding=pad, gain=np.sqrt(2), slope=0.2, clamp=clamp)
    assert y is not None
    y = y.cpu().numpy()
    assert y.shape == ref.shape
    err = np.abs(y - ref).max() / np.abs(ref).max()
    assert err <= TOL, err


def test_tc_odd_width_is_unsupported():
    """The tensor-core kernel reads aligned column pairs; an odd width reports 'unsupported' (None) and the
    caller uses the exact kernel, like the reference's return code -1 (OPS/filtered_lrelu.cpp:52-56)."""
    import torch
    from afcm_b200.torch_utils.ops.filtered_lrelu import filtered_lrelu_tc
    dev = torch.device('cuda')
    fu, fd = _filters(2, 2)
    x = torch.randn(1, 2, 20, 21, device=dev)
    y = filtered_lrelu_tc(x, torch.from_numpy(fu).to(dev), torch.from_numpy(fd).to(dev), None, up=2, down=2,
                          padding=[9, 8, 9, 8], clamp=256.0)
    assert y is None


def test_tc_skip_scale_fp16_io():
    import torch
    from afcm_b200.torch_utils.ops.filtered_lrelu import filtered_lrelu_tc
    from oracle import afcm_oracle as orc
    rng = np.random.RandomState(5)
    up, down, pad = 2, 2, [9, 8, 9, 8]
    fu, fd = _filters(up, down)
    x = rng.randn(2, 4, 54, 54).astype(np.float32)
    b = rng.randn(4).astype(np.float32)
    ref = orc.filtered_lrelu(x, fu, fd, b, up=up, down=down, padding=pad, gain=np.sqrt(2), slope=0.2, clamp=256.0)
    ref = ref[0] if isinstance(ref, tuple) else ref
    skip = rng.randn(*ref.shape).astype(np.float32)
    want = (ref + skip) * 0.25
    dev = torch.device('cuda')
    t = lambda a: torch.from_numpy(a).to(dev)
    y = filtered_lrelu_tc(t(x), t(fu), t(fd), t(b), up=up, down=down, padding=pad, gain=np.sqrt(2), slope=0.2, clamp=256.0,
                          skip=t(skip), out_scale=0.25)
    assert np.abs(y.cpu().numpy() - want).max() <= TOL * np.abs(want).max()
    # fp16 in / fp16 out
    y16 = filtered_lrelu_tc(t(x).half(), t(fu), t(fd), t(b), up=up, down=down, padding=pad, gain=np.sqrt(2), slope=0.2,
                            clamp=256.0, out_dtype=torch.float16)
    assert y16.dtype == torch.float16
    assert np.abs(y16.float().cpu().numpy() - ref).max() <= 2 * TOL * np.abs(ref).max()
    # padded row pitch (a view into a wider buffer) is honoured
    buf = torch.zeros(2, 4, 54, 56, device=dev)
    buf[..., :54] = t(x)
    y2 = filtered_lrelu_tc(buf[..., :54], t(fu), t(fd), t(b), up=up, down=down, padding=pad, gain=np.sqrt(2), slope=0.2,
                           clamp=256.0)
    assert np.abs(y2.cpu().numpy() - ref).max() <= TOL * np.abs(ref).max()


def test_tc_matches_exact_kernel_full_size():
    """BASELINE.json-sized planes (278 -> 276, 278 -> 148, 150 -> 276) against the exact fp32 kernel."""
    import torch
    from afcm_b200.networks_stylegan3 import afcm_generator
    from afcm_b200.torch_utils.ops.filtered_lrelu import _run_fused, filtered_lrelu_tc
    dev = torch.device('cuda')
    G = afcm_generator(seed=0, device=dev)
    S = G.synthesis
    torch.manual_seed(0)
    for name in ['encoder_1', 'encoder_4', 'L10_276_128', 'L13_256_64']:
        L = getattr(S, name)
        Hc = int(L.in_size[0]) + 2
        x = torch.randn(2, L.out_channels, Hc, Hc, device=dev) * 3
        b = torch.randn(L.out_channels, device=dev)
        px0, px1, py0, py1 = L.padding
        ref, _, rc = _run_fused(x, L.up_filter, L.down_filter, b, None, L.up_factor, L.down_factor, px0, px1, py0, py1, 0, 0,
                                float(np.sqrt(2)), 0.2, 256.0, False, False)
        assert rc == 0
        y = filtered_lrelu_tc(x, L.up_filter, L.down_filter, b, up=L.up_factor, down=L.down_factor, padding=L.padding,
                              gain=np.sqrt(2), slope=0.2, clamp=256.0)
        assert y is not None and y.shape == ref.shape
        err = float((y - ref).abs().max() / ref.abs().max())
        assert err <= TOL, (name, err)


@pytest.mark.parametrize('up,down,pad,N,C,H,W', CASES)
def test_tc_fast_variant_matches_oracle(up, down, pad, N, C, H, W):
    """The variant the fast inference path runs: fp16 planes in and out, no bias (the convolution epilogue added it),
    no skip tensor, sat() activation."""
    import torch
    from afcm_b200.torch_utils.ops.filtered_lrelu import filtered_lrelu_tc
    from oracle import afcm_oracle as orc
    rng = np.random.RandomState(H * 11 + W)
    fu, fd = _filters(up, down)
    x = (rng.randn(N, C, H, W) * 2).astype(np.float16)
    ref = orc.filtered_lrelu(x.astype(np.float32), fu, fd, None, up=up, down=down, padding=pad, gain=np.sqrt(2), slope=0.2, clamp=3.0)
    ref = ref[0] if isinstance(ref, tuple) else ref
    dev = torch.device('cuda')
    y = filtered_lrelu_tc(torch.from_numpy(x).to(dev), torch.from_numpy(fu).to(dev), torch.from_numpy(fd).to(dev), None,
                          up=up, down=down, padding=pad, gain=np.sqrt(2), slope=0.2, clamp=3.0, out_dtype=torch.float16)
    assert y is not None and y.dtype == torch.float16
    y = y.float().cpu().numpy()
    assert y.shape == ref.shape
    err = np.abs(y - ref).max() / np.abs(ref).max()
    assert err <= 2 * TOL, err


FULL_CASES = [  # the plane sizes that carry most bytes of the BASELINE forward: up, down, padding, H
    (2, 2, [9, 8, 9, 8], 278),            # 278 -> 276
    (2, 4, [34, 33, 34, 33], 278),        # 278 -> 148
    (4, 2, [-6, -9, -6, -9], 150),        # 150 -> 276
    (2, 2, [-11, -12, -11, -12], 278),    # 278 -> 256 (last synthesis layer)
    (2, 4, [34, 33, 34, 33], 150),        # 150 -> 84
    (4, 2, [-6, -9, -6, -9], 86),         # 86 -> 148
]


@pytest.mark.parametrize('up,down,pad,H', FULL_CASES)
def test_tc_fast_variant_full_size_vs_oracle(up, down, pad, H):
    """Per-operator parity of the benchmarked variant (fp16 planes in and out, no bias, sat() activation, clamp 256) at the
    full-size geometries, against the CPU oracle: max|err| <= 2e-3 max|y| (fp16 operands + fp16 result rounding)."""
    import torch
    from afcm_b200.torch_utils.ops.filtered_lrelu import filtered_lrelu_tc
    from oracle import afcm_oracle as orc
    rng = np.random.RandomState(H + 17 * up + down)
    fu, fd = _filters(up, down)
    x = (rng.randn(1, 3, H, H) * 3).astype(np.float16)
    ref = orc.filtered_lrelu(x.astype(np.float32), fu, fd, None, up=up, down=down, padding=pad, gain=np.sqrt(2), slope=0.2, clamp=256.0)
    ref = ref[0] if isinstance(ref, tuple) else ref
    dev = torch.device('cuda')
    y = filtered_lrelu_tc(torch.from_numpy(x).to(dev), torch.from_numpy(fu).to(dev), torch.from_numpy(fd).to(dev), None,
                          up=up, down=down, padding=pad, gain=np.sqrt(2), slope=0.2, clamp=256.0, out_dtype=torch.float16)
    assert y is not None and y.dtype == torch.float16 and tuple(y.shape) == ref.shape
    err = np.abs(y.float().cpu().numpy() - ref).max() / np.abs(ref).max()
    print(f'flr_tc fast variant {H}px up{up} down{down}: rel err {err:.3e}')
    assert err <= TOL, err
