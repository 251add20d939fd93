"""GPU parity of the tcgen05 / TMEM filtered_lrelu (afcm_filtered_lrelu_t5, csrc/flr_t5.cu) against the CPU oracle at every
AFCM geometry incl. the full-size planes.  Tolerance (fp16 operands, tf32 horizontal up pass, fp32 accumulation, fp16 result):
max|err| <= 2e-3 * max|y| per call -- the same bound as the mma.sync kernel (tests/test_gpu_flr_tc.py)."""
import numpy as np
import pytest
import scipy.signal

pytestmark = pytest.mark.gpu
TOL = 2e-3

CASES = [  # up, down, padding, N, C, H, W, skip
    (2, 2, [9, 8, 9, 8], 1, 1, 22, 26, False),
    (2, 2, [9, 8, 9, 8], 2, 3, 38, 38, False),
    (2, 2, [9, 8, 9, 8], 2, 3, 38, 38, True),
    (2, 2, [9, 8, 9, 8], 1, 2, 70, 150, False),
    (2, 2, [-11, -12, -11, -12], 1, 3, 38, 36, False),
    (2, 2, [9, 8, 7, 10], 2, 2, 21, 20, False),
    (2, 2, [8, 9, 10, 7], 1, 2, 70, 84, False),
    (4, 2, [-6, -9, -6, -9], 2, 3, 22, 26, False),
    (4, 2, [-6, -9, -6, -9], 1, 4, 38, 38, True),
    (4, 2, [3, 2, 1, 4], 1, 3, 9, 12, False),
    (4, 2, [-5, -10, -7, -8], 1, 2, 54, 54, False),
    (2, 4, [34, 33, 34, 33], 2, 3, 38, 42, False),
    (2, 4, [34, 33, 34, 33], 1, 2, 54, 54, True),
    (2, 4, [33, 34, 35, 32], 1, 2, 86, 86, False),
    (2, 2, [9, 8, 9, 8], 1, 2, 278, 278, False),
    (2, 4, [34, 33, 34, 33], 1, 2, 278, 278, False),
    (4, 2, [-6, -9, -6, -9], 1, 2, 150, 150, False),
    (2, 2, [-11, -12, -11, -12], 1, 2, 278, 278, False),
]


@pytest.mark.parametrize('up,down,pad,N,C,H,W,skip', CASES)
def test_t5_matches_oracle(up, down, pad, N, C, H, W, skip):
    import torch
    from afcm_b200.torch_utils.ops import filtered_lrelu as flr
    from oracle import afcm_oracle as orc
    dev = torch.device('cuda')
    rng = np.random.RandomState(H * 11 + W)
    fu = scipy.signal.firwin(6 * up, 0.4, width=0.3, fs=2).astype(np.float32)
    fd = scipy.signal.firwin(6 * down, 0.25, width=0.2, fs=2).astype(np.float32)
    x = (rng.randn(N, C, H, W) * 2).astype(np.float16)
    clamp = 256.0 if skip else 3.0
    ref = orc.filtered_lrelu(x.astype(np.float32), fu, fd, None, up=up, down=down, padding=pad, gain=np.sqrt(2), slope=0.2, clamp=clamp)
    ref = ref[0] if isinstance(ref, tuple) else ref
    sk, scale = None, 1.0
    if skip:
        sk = rng.randn(*ref.shape).astype(np.float16)
        scale = 0.25
        ref = (ref + sk.astype(np.float32)) * scale
    xv = flr.padded_pitch_empty([N, C, H, W], torch.float16, dev)        # TMA needs a 16-byte aligned row pitch
    xv.copy_(torch.from_numpy(x).to(dev))
    t = lambda a: torch.from_numpy(a).to(dev)
    y = flr.filtered_lrelu_tc(xv, t(fu), t(fd), None, up=up, down=down, padding=pad, gain=np.sqrt(2), slope=0.2, clamp=clamp,
                              out_dtype=torch.float16, skip=None if sk is None else t(sk), out_scale=scale, impl='t5')
    assert y is not None, flr._lib.last_error()
    err = np.abs(y.float().cpu().numpy() - ref).max() / np.abs(ref).max()
    assert err <= TOL, err


def test_t5_reports_unsupported_like_the_reference():
    """A call the kernel has no specialisation for (bias given, unaligned pitch) answers 'unsupported' and the caller falls back,
    like return code -1 of the reference plugin (OPS/filtered_lrelu.cpp:52-56)."""
    import torch
    from afcm_b200.torch_utils.ops import filtered_lrelu as flr
    dev = torch.device('cuda')
    fu = torch.from_numpy(scipy.signal.firwin(12, 0.4, width=0.3, fs=2).astype(np.float32)).to(dev)
    x = torch.randn(1, 2, 38, 38, device=dev, dtype=torch.float16)      # row pitch 76 bytes: not a multiple of 16
    y = flr.filtered_lrelu_tc(x, fu, fu, None, up=2, down=2, padding=[9, 8, 9, 8], clamp=256.0, impl='t5')
    assert y is None
