"""CPU check of the index math of the tensor-core filtered_lrelu (csrc/flr_tc.cu): the numpy model of the
strip algorithm (tools/flr_tc_model.py: same phase shift / alignment / origin formulas and the same closed-form
Toeplitz operators the kernel builds) must reproduce the oracle."""
import numpy as np
import pytest
import scipy.signal

from oracle import afcm_oracle as orc
from tools.flr_tc_model import filtered_lrelu_model

CASES = [  # up, down, padding, H, W
    (2, 2, [9, 8, 9, 8], 22, 25),          # encoder / synthesis same-resolution layers
    (2, 4, [34, 33, 34, 33], 38, 41),      # encoder down-sampling layers
    (4, 2, [-6, -9, -6, -9], 22, 25),      # synthesis up-sampling layers (cropping)
    (2, 2, [-11, -12, -11, -12], 38, 35),  # L13 (crop to 256)
    (2, 2, [9, 8, 7, 10], 21, 20),         # asymmetric / odd
    (4, 2, [3, 2, 1, 4], 9, 12),
]


@pytest.mark.parametrize('up,down,pad,H,W', CASES)
def test_model_matches_oracle(up, down, pad, H, W):
    rng = np.random.RandomState(up * 10 + down)
    fu = scipy.signal.firwin(6 * up, 0.4, width=0.3, fs=2).astype(np.float32)
    fd = scipy.signal.firwin(6 * down, 0.25, width=0.2, fs=2).astype(np.float32)
    x = (rng.randn(2, 3, H, W) * 2).astype(np.float32)
    b = rng.randn(3).astype(np.float32)
    ref = orc.filtered_lrelu(x, fu, fd, b, up=up, down=down, padding=pad, gain=np.sqrt(2), slope=0.2, clamp=2.0)
    ref = ref[0] if isinstance(ref, tuple) else ref
    y = filtered_lrelu_model(x, fu, fd, b, up, down, pad, np.sqrt(2), 0.2, 2.0)
    assert y.shape == ref.shape
    assert np.abs(y - ref).max() <= 1e-5 * np.abs(ref).max()
    y16 = filtered_lrelu_model(x, fu, fd, b, up, down, pad, np.sqrt(2), 0.2, 2.0, fp16=True)
    assert np.abs(y16 - ref).max() <= 2e-3 * np.abs(ref).max()       # the tolerance stated for the fp16 path
