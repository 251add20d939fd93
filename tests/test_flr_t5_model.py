"""CPU check of the design model of a tcgen05/TMEM filtered_lrelu (tools/flr_t5_model.py): the blocked MN-major pass chain
with only the in-band Toeplitz blocks issued reproduces the oracle, with fp16 operand rounding it stays inside the 2e-3
tolerance of the tensor-core path, and the cost model gives the numbers DESIGN.md section 7 quotes."""
import numpy as np
import pytest
import scipy.signal

from oracle import afcm_oracle as orc
from tools.flr_t5_model import filtered_lrelu_t5_model

CASES = [  # up, down, padding, H, W, tile, NB
    (2, 2, [9, 8, 9, 8], 70, 66, (56, 56), (64, 64, 32, 32)),
    (2, 2, [8, 9, 10, 7], 37, 52, (24, 40), (32, 32, 16, 16)),       # odd phase shift, ragged tiles
    (2, 4, [34, 33, 34, 33], 86, 86, (24, 24), (64, 64, 16, 16)),
    (4, 2, [-6, -9, -6, -9], 38, 38, (56, 56), (64, 64, 32, 32)),
    (2, 2, [-11, -12, -11, -12], 38, 36, (16, 16), (16, 16, 16, 16)),
]


@pytest.mark.parametrize('up,down,pad,H,W,tile,NB', CASES)
def test_t5_model_matches_oracle(up, down, pad, H, W, tile, NB):
    rng = np.random.RandomState(up * 10 + down + H)
    fu = scipy.signal.firwin(6 * up, 0.4, width=0.3, fs=2).astype(np.float32)
    fd = scipy.signal.firwin(6 * down, 0.25, width=0.2, fs=2).astype(np.float32)
    x = (rng.randn(1, 2, H, W) * 2).astype(np.float32)
    b = rng.randn(2).astype(np.float32)
    ref = orc.filtered_lrelu(x, fu, fd, b, up=up, down=down, padding=pad, gain=np.sqrt(2), slope=0.2, clamp=2.0)
    ref = ref[0] if isinstance(ref, tuple) else ref
    y, _ = filtered_lrelu_t5_model(x, fu, fd, b, up, down, pad, np.sqrt(2), 0.2, 2.0, tile=tile, NB=NB)
    assert y.shape == ref.shape
    assert np.abs(y - ref).max() <= 1e-5 * np.abs(ref).max()
    y16, _ = filtered_lrelu_t5_model(x, fu, fd, b, up, down, pad, np.sqrt(2), 0.2, 2.0, fp16=True, tile=tile, NB=NB)
    assert np.abs(y16 - ref).max() <= 2e-3 * np.abs(ref).max()


def test_t5_cost_model_is_shared_memory_bound():
    """The finding recorded in DESIGN.md: every M = 128 instruction re-reads its 4 KB A slice from shared memory, so with
    the narrow in-band blocks the plan is shared-memory bound, about 2x its tensor time, and slower than the HBM rate."""
    fu = scipy.signal.firwin(12, 0.4, width=0.3, fs=2).astype(np.float32)
    x = np.zeros((1, 1, 240, 240), np.float32)
    _, st = filtered_lrelu_t5_model(x, fu, fu, None, 2, 2, [9, 8, 9, 8], np.sqrt(2), 0.2, 2.0, tile=(112, 112), NB=(64, 64, 32, 32))
    tensor = sum(v['tensor_clk'] for v in st['passes'].values()) * 256 / (112 * 112)
    assert 70 < st['mma_clk_per_256_outputs'] < 85 and 35 < tensor < 45
    assert st['mma_clk_per_256_outputs'] > 1.8 * st['hbm_clk_per_256_outputs_fp16']
    assert st['epilogue_clk_per_256_outputs'] < st['hbm_clk_per_256_outputs_fp16']
