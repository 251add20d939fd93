"""CPU check of the index algebra of the tcgen05 / TMEM filtered_lrelu (csrc/flr_t5.cu): the numpy emulation of its
data flow (tools/flr_t5_emu.py: per-step Toeplitz products, groups of 64 up-sampled columns, first-touch order of the
overlapping accumulator windows, the shared-memory ring between the horizontal and the vertical down pass) must reproduce
the oracle, with the PLAN taken from the library's own host code (afcm_filtered_lrelu_t5_plan -> flr_t5_plan.h)."""
import ctypes

import numpy as np
import pytest
import scipy.signal

from afcm_b200 import _lib
from oracle import afcm_oracle as orc
from tools.flr_t5_emu import filtered_lrelu_t5_emu, t5_plan_py

PLAN_FIELDS = ['U', 'D', 'FU', 'FD', 'xh', 'xw', 'yh', 'yw', 'RS', 'K1', 'I0y', 'tuy_e', 'OS', 'wlo0', 'nsteps', 'NL', 't4_e',
               'KW', 'nstrips', 'iorg0', 'istep', 'jorg0', 'korg0', 'm0', 't3_e', 'NG', 'N1', 'halves']


def lib_plan(xh, xw, up, down, pad, kw=0):
    out = (ctypes.c_int * 64)()
    rc = _lib.lib().afcm_filtered_lrelu_t5_plan(xh, xw, up, down, pad[0], pad[1], pad[2], pad[3], kw, out, 64)
    assert rc == 0, _lib.last_error()
    p = {k: int(out[i]) for i, k in enumerate(PLAN_FIELDS)}
    p['px0'], p['py0'] = pad[0], pad[2]
    p['adv3'] = p['adv4'] = 16 // down
    return p


CASES = [  # up, down, padding, H, W, kw (0 = automatic)
    (2, 2, [9, 8, 9, 8], 22, 25, 0),           # same-resolution layers
    (2, 4, [34, 33, 34, 33], 38, 41, 0),       # encoder down-sampling layers
    (4, 2, [-6, -9, -6, -9], 22, 25, 0),       # synthesis up-sampling layers (cropping)
    (2, 2, [-11, -12, -11, -12], 38, 35, 0),   # L13 (crop to 256)
    (2, 2, [9, 8, 7, 10], 21, 20, 0),          # asymmetric / odd
    (4, 2, [3, 2, 1, 4], 9, 12, 0),
    (2, 2, [9, 8, 9, 8], 70, 150, 0),          # two strips, two steps
    (2, 2, [9, 8, 9, 8], 40, 278, 96),         # three strips of the 276-pixel geometry, four groups
    (2, 4, [34, 33, 34, 33], 86, 150, 0),      # two strips, lead chunks of the 24-tap filter across steps
    (4, 2, [-6, -9, -6, -9], 54, 86, 0),       # two strips, three steps
]


@pytest.mark.parametrize('up,down,pad,H,W,kw', CASES)
def test_plan_matches_python_statement(up, down, pad, H, W, kw):
    a = lib_plan(H, W, up, down, pad, kw)
    b = t5_plan_py(H, W, up, down, pad, kw=kw or None)
    for k in PLAN_FIELDS:
        if k in b:
            assert a[k] == b[k], k
    assert a['N1'] <= 128 and a['m0'] + a['KW'] <= 128 and a['iorg0'] % 8 == 0 and a['istep'] % 8 == 0


@pytest.mark.parametrize('up,down,pad,H,W,kw', CASES)
def test_emulation_matches_oracle(up, down, pad, H, W, kw):
    rng = np.random.RandomState(up * 10 + down)
    fu = scipy.signal.firwin(6 * up, 0.4, width=0.3, fs=2).astype(np.float32)
    fd = scipy.signal.firwin(6 * down, 0.25, width=0.2, fs=2).astype(np.float32)
    x = (rng.randn(1, 2, H, W) * 2).astype(np.float32)
    ref = orc.filtered_lrelu(x, fu, fd, None, up=up, down=down, padding=pad, gain=np.sqrt(2), slope=0.2, clamp=2.0)
    ref = ref[0] if isinstance(ref, tuple) else ref
    plan = lib_plan(H, W, up, down, pad, kw)
    y = filtered_lrelu_t5_emu(x, fu, fd, up, down, pad, np.sqrt(2), 0.2, 2.0, plan=plan)
    assert y.shape == ref.shape
    assert np.abs(y - ref).max() <= 1e-5 * np.abs(ref).max()
    y16 = filtered_lrelu_t5_emu(x, fu, fd, up, down, pad, np.sqrt(2), 0.2, 2.0, plan=plan, rounded=True)
    assert np.abs(y16 - ref).max() <= 2e-3 * np.abs(ref).max()      # the tolerance stated for the kernel


def test_afcm_layer_plans_fit():
    """Every filtered_lrelu geometry of the AFCM generator has a plan: strips, steps, TMEM columns."""
    geos = [(2, 2, [9, 8, 9, 8], 278), (2, 4, [34, 33, 34, 33], 278), (2, 2, [9, 8, 9, 8], 150), (2, 4, [34, 33, 34, 33], 150),
            (2, 2, [9, 8, 9, 8], 86), (2, 4, [34, 33, 34, 33], 86), (2, 2, [9, 8, 9, 8], 54), (2, 4, [34, 33, 34, 33], 54),
            (2, 2, [9, 8, 9, 8], 38), (4, 2, [-6, -9, -6, -9], 38), (4, 2, [-6, -9, -6, -9], 54), (4, 2, [-6, -9, -6, -9], 86),
            (4, 2, [-6, -9, -6, -9], 150), (2, 2, [-11, -12, -11, -12], 278)]
    for up, down, pad, n in geos:
        p = lib_plan(n, n, up, down, pad)
        assert p['nstrips'] * p['KW'] >= p['yw'] and p['NG'] * 64 // up == p['N1']
        assert 40 + 128 + 128 + (16 // down) * (4 * p['NG'] - 1) + 16 <= 432        # D3 ends where D4 starts (flr_t5.cu TMEM map)
