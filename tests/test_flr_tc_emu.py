"""CPU check of the schedule of the tensor-core filtered_lrelu (csrc/flr_tc.cu) through its block-level emulation
(tools/flr_tc_emu.py): the same state machine -- row-block ring, P slots, carry, super-iterations with the pipeline fill /
drain skips, narrow last strip, row segments -- built from the same index expressions reproduces the oracle at the
geometries of the GPU parity test, exactly in float64 and within the stated 2e-3 with the kernel's fp16 rounding points."""
import numpy as np
import pytest
import scipy.signal

from oracle import afcm_oracle as orc
from tools.flr_tc_emu import filtered_lrelu_tc_emu

CASES = [  # up, down, padding, H, W, seg_wblocks  (tests/test_gpu_flr_tc.py CASES + segmented strips)
    (2, 2, [9, 8, 9, 8], 22, 26, None),
    (2, 2, [9, 8, 9, 8], 38, 38, None),
    (2, 4, [34, 33, 34, 33], 38, 42, None),
    (2, 4, [34, 33, 34, 33], 54, 54, None),
    (4, 2, [-6, -9, -6, -9], 22, 26, None),
    (4, 2, [-6, -9, -6, -9], 38, 38, None),
    (2, 2, [-11, -12, -11, -12], 38, 36, None),
    (2, 2, [9, 8, 7, 10], 21, 20, None),
    (4, 2, [3, 2, 1, 4], 9, 12, None),
    (2, 2, [8, 9, 10, 7], 70, 84, None),
    (2, 4, [33, 34, 35, 32], 86, 86, None),
    (4, 2, [-5, -10, -7, -8], 54, 54, None),
    (2, 2, [9, 8, 9, 8], 86, 38, 4),           # row segments of 32 output rows
    (2, 4, [34, 33, 34, 33], 86, 54, 4),
    (4, 2, [-6, -9, -6, -9], 38, 22, 4),
]


@pytest.mark.parametrize('up,down,pad,H,W,seg', CASES)
def test_emulation_matches_oracle(up, down, pad, H, W, seg):
    rng = np.random.RandomState(H * 7 + W)
    fu = scipy.signal.firwin(6 * up, 0.4, width=0.3, fs=2).astype(np.float32)
    fd = scipy.signal.firwin(6 * down, 0.25, width=0.2, fs=2).astype(np.float32)
    x = (rng.randn(1, 2, H, W) * 2).astype(np.float32)
    b = rng.randn(2).astype(np.float32)
    ref = orc.filtered_lrelu(x, fu, fd, b, up=up, down=down, padding=pad, gain=np.sqrt(2), slope=0.2, clamp=2.0)
    ref = ref[0] if isinstance(ref, tuple) else ref
    y = filtered_lrelu_tc_emu(x, fu, fd, b, up, down, pad, np.sqrt(2), 0.2, 2.0, seg_wblocks=seg)
    assert y.shape == ref.shape
    assert np.abs(y - ref).max() <= 1e-5 * np.abs(ref).max()
    y16 = filtered_lrelu_tc_emu(x, fu, fd, b, up, down, pad, np.sqrt(2), 0.2, 2.0, fp16=True, seg_wblocks=seg)
    assert np.abs(y16 - ref).max() <= 2e-3 * np.abs(ref).max()


@pytest.mark.parametrize('up,down,pad,H,W,seg', [(2, 2, [9, 8, 9, 8], 38, 38, None), (2, 2, [8, 9, 10, 7], 38, 52, 4),
                                                 (2, 4, [34, 33, 34, 33], 54, 54, None), (4, 2, [-6, -9, -6, -9], 22, 26, None)])
def test_fragment_coordinates_of_the_sign_tensor(up, down, pad, H, W, seg):
    """Groundwork for a sign mode of the kernel: the pre-activation value a warp holds in fragment element (column block mb,
    row m; row block nb, column n) is the up-sampled sample the reference sign tensor indexes at (uy, ux) as stated in
    Warp.record -- every sample of the sign tensor's extent is produced by some warp and equals the oracle's value."""
    rng = np.random.RandomState(H + W)
    fu = scipy.signal.firwin(6 * up, 0.4, width=0.3, fs=2).astype(np.float32)
    fd = scipy.signal.firwin(6 * down, 0.25, width=0.2, fs=2).astype(np.float32)
    x = (rng.randn(1, 1, H, W) * 2).astype(np.float32)
    _, so, pre = orc.filtered_lrelu(x, fu, fd, None, up=up, down=down, padding=pad, gain=np.sqrt(2), slope=0.2, clamp=2.0,
                                    write_signs=True, return_preact=True)
    sz = orc.filtered_lrelu_sizes(H, W, up, down, len(fu), len(fd), pad)
    sh = sz['SH']
    sw = sz['OW'] * down - (down - 1) + len(fd) - 1                      # written width of the sign tensor (its pitch is padded)
    _, got = filtered_lrelu_tc_emu(x, fu, fd, None, up, down, pad, np.sqrt(2), 0.2, 2.0, seg_wblocks=seg, preact_shape=(sh, sw))
    assert not np.isnan(got).any()
    ref = pre[:, :, :sh, :sw]
    assert np.abs(got - ref).max() <= 1e-5 * np.abs(ref).max()


@pytest.mark.parametrize('up,down,pad,H,W,seg', [(2, 2, [9, 8, 9, 8], 38, 38, None), (2, 2, [8, 9, 10, 7], 38, 52, 4),
                                                 (2, 2, [9, 8, 9, 8], 22, 26, None), (2, 4, [34, 33, 34, 33], 54, 54, None),
                                                 (4, 2, [-6, -9, -6, -9], 22, 26, None), (4, 2, [-5, -10, -7, -8], 54, 38, 4)])
def test_sign_write_mode_design(up, down, pad, H, W, seg):
    """The planned sign-write mode, emulated: codes taken from the fragments each warp holds, packed with the warp recipe,
    stored under the ownership rule of Warp.flush_signs, give the oracle's sign tensor bit for bit (columns the reference
    never reads -- beyond the written width -- excluded)."""
    rng = np.random.RandomState(H * 3 + W)
    fu = scipy.signal.firwin(6 * up, 0.4, width=0.3, fs=2).astype(np.float32)
    fd = scipy.signal.firwin(6 * down, 0.25, width=0.2, fs=2).astype(np.float32)
    x = (rng.randn(1, 2, H, W) * 2).astype(np.float32)
    b = rng.randn(2).astype(np.float32)
    _, so = orc.filtered_lrelu(x, fu, fd, b, up=up, down=down, padding=pad, gain=np.sqrt(2), slope=0.2, clamp=1.0, write_signs=True)
    sz = orc.filtered_lrelu_sizes(H, W, up, down, len(fu), len(fd), pad)
    sw = sz['OW'] * down - (down - 1) + len(fd) - 1
    _, got = filtered_lrelu_tc_emu(x, fu, fd, b, up, down, pad, np.sqrt(2), 0.2, 1.0, seg_wblocks=seg, sign_shape=so.shape[2:])
    unpack = lambda s: np.stack([(s >> (2 * j)) & 3 for j in range(4)], -1).reshape(*s.shape[:3], -1)[..., :sw]
    assert np.array_equal(unpack(got), unpack(so))


@pytest.mark.parametrize('up,down,pad,H,W', [(2, 2, [9, 8, 9, 8], 38, 38), (2, 2, [8, 9, 10, 7], 38, 52), (2, 4, [34, 33, 34, 33], 54, 54),
                                             (4, 2, [-6, -9, -6, -9], 22, 26)])
def test_sign_read_mode_design(up, down, pad, H, W):
    """The backward pass as the planned sign-read mode: the op with up / down and the filters exchanged, the padding, gain
    and sign offsets of OPS/filtered_lrelu.py:252-263, reading the forward's sign tensor at the fragment coordinates."""
    rng = np.random.RandomState(H * 5 + W)
    fu = scipy.signal.firwin(6 * up, 0.4, width=0.3, fs=2).astype(np.float32)
    fd = scipy.signal.firwin(6 * down, 0.25, width=0.2, fs=2).astype(np.float32)
    x = (rng.randn(1, 2, H, W) * 2).astype(np.float32)
    y, so = orc.filtered_lrelu(x, fu, fd, None, up=up, down=down, padding=pad, gain=np.sqrt(2), slope=0.2, clamp=1.0, write_signs=True)
    dy = rng.randn(*y.shape).astype(np.float32)
    px0, px1, py0, py1 = pad
    yh, yw = y.shape[2:]
    pp = [(len(fu) - 1) + (len(fd) - 1) - px0, W * up - yw * down + px0 - (up - 1),
          (len(fu) - 1) + (len(fd) - 1) - py0, H * up - yh * down + py0 - (up - 1)]
    gg = np.sqrt(2) * up ** 2 / down ** 2
    sxo, syo = -(len(fu) - 1) + px0, -(len(fu) - 1) + py0
    ref = orc.filtered_lrelu(dy, fd, fu, None, up=down, down=up, padding=pp, gain=gg, slope=0.2, clamp=None, flip_filter=True,
                             si=so, sx=sxo, sy=syo)
    assert ref.shape == x.shape
    got = filtered_lrelu_tc_emu(dy, fd, fu, None, down, up, pp, gg, 0.2, None, flip_filter=True, si=so, s_ofs=(sxo, syo))
    assert np.abs(got - ref).max() <= 1e-5 * np.abs(ref).max()
