"""CPU emulation of the fused filtered_lrelu kernel passes (tests/emu/flr_emu.cpp compiles the very
same __host__ __device__ pass functions the CUDA kernel runs) checked against the oracle and the
reference golden vectors.  Catches tile / polyphase / sign indexing bugs without a GPU."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import rel_err, ROOT
from oracle import afcm_oracle as orc

EMU_DIR = os.path.join(ROOT, 'tests', 'emu')


@pytest.fixture(scope='module')
def emu():
    so = os.path.join(EMU_DIR, 'libflr_emu.so')
    srcs = [os.path.join(EMU_DIR, 'flr_emu.cpp'), os.path.join(ROOT, 'afcm_b200', 'csrc', 'filtered_lrelu_core.h')]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(['g++', '-O1', '-shared', '-fPIC', '-std=c++17', '-ffp-contract=off', '-o', so, srcs[0]])
    return ctypes.CDLL(so)


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def run_emu(emu, x, fu, fd, b, up, down, pad, gain, slope, clamp, flip=False, sign_mode=0, signs=None, sx=0, sy=0,
            tile=(16, 8), nthr=64, skip=None, out_scale=1.0):
    x = np.ascontiguousarray(x, np.float32)
    N, C, xh, xw = x.shape
    sz = orc.filtered_lrelu_sizes(xh, xw, up, down, len(fu), len(fd), pad)
    y = np.full([N, C, sz['OH'], sz['OW']], np.nan, np.float32)
    sh = swb = 0
    if sign_mode == 1:
        signs = np.full([N, C, sz['SH'], sz['SWB']], 0xAA, np.uint8)
    if signs is not None:
        sh, swb = signs.shape[2:]
    cl = np.float32(np.inf if clamp is None else clamp)
    rc = emu.emu_filtered_lrelu(_p(x), _p(y), _p(b), _p(skip), N, C, xh, xw, sz['OH'], sz['OW'],
                                _p(np.asarray(fu, np.float32)), len(fu), _p(np.asarray(fd, np.float32)), len(fd),
                                up, down, pad[0], pad[2], ctypes.c_float(gain), ctypes.c_float(slope), ctypes.c_float(cl),
                                ctypes.c_float(out_scale), int(flip), sign_mode, _p(signs), sh, swb, sx, sy,
                                tile[0], tile[1], nthr)
    assert rc == 0
    return (y, signs) if sign_mode == 1 else y


def _cases(g):
    for name in g['flrelu.names']:
        k = 'flrelu.' + str(name)
        c = g[k + '.cfg']
        fu, fd = g[k + '.fu'], g[k + '.fd']
        if fu.size not in (12, 24) or fd.size not in (12, 24):
            continue
        yield str(name), k, int(c[0]), int(c[1]), [int(v) for v in c[2:6]], float(c[6]), float(c[7]), \
            (None if c[8] < 0 else float(c[8])), fu, fd


@pytest.mark.parametrize('tile', [(16, 8), (24, 16), (8, 8), (64, 32)])
def test_emulated_forward_matches_reference_golden(emu, golden_ops, tile):
    g = golden_ops
    n = 0
    for name, k, up, dn, pad, gain, slope, clamp, fu, fd in _cases(g):
        y = run_emu(emu, g[k + '.x'], fu, fd, g[k + '.b'], up, dn, pad, gain, slope, clamp, tile=tile)
        assert not np.isnan(y).any(), name
        assert rel_err(y, g[k + '.y']) < 3e-6, (name, tile)
        yf = run_emu(emu, g[k + '.x'], fu, fd, g[k + '.b'], up, dn, pad, gain, slope, clamp, flip=True, tile=tile)
        assert rel_err(yf, g[k + '.yflip']) < 3e-6, (name, tile)
        n += 1
    assert n >= 6


def test_emulated_signs_and_backward(emu, golden_ops):
    g = golden_ops
    for name, k, up, dn, pad, gain, slope, clamp, fu, fd in _cases(g):
        x = g[k + '.x']
        y, so = run_emu(emu, x, fu, fd, g[k + '.b'], up, dn, pad, gain, slope, clamp, sign_mode=1, tile=(16, 8))
        yo, so_o, pre = orc.filtered_lrelu(x, fu, fd, g[k + '.b'], up, dn, pad, gain, slope, clamp, write_signs=True,
                                           return_preact=True)
        assert rel_err(y, g[k + '.y']) < 3e-6, name
        # compare sign codes on the active region, ignoring elements whose pre-activation is within
        # rounding distance of a decision threshold (0 or +-clamp)
        sz = orc.filtered_lrelu_sizes(x.shape[2], x.shape[3], up, dn, len(fu), len(fd), pad)
        sw_active = sz['OW'] * dn - (dn - 1) + len(fd) - 1
        def unpack(s):
            e = np.stack([(s >> (2 * j)) & 3 for j in range(4)], -1).reshape(*s.shape[:3], -1)
            return e[..., :sw_active]
        a, b_ = unpack(so), unpack(so_o)
        pre = pre[:, :, :sz['SH'], :sw_active]
        cl = np.inf if clamp is None else clamp
        act = np.where(pre < 0, pre * slope, pre)
        safe = ((np.abs(pre) > 1e-5) | (pre == 0)) & (np.abs(np.abs(act) - cl) > 1e-4 * max(cl, 1) if np.isfinite(cl) else True)
        assert (a[safe] == b_[safe]).all(), name
        assert safe.mean() > 0.98
        # backward = same op, swapped filters, reading signs (OPS/filtered_lrelu.py:252-266)
        xh, xw = x.shape[2:]; yh, yw = y.shape[2:]
        pp = [(len(fu) - 1) + (len(fd) - 1) - pad[0], xw * up - yw * dn + pad[0] - (up - 1),
              (len(fu) - 1) + (len(fd) - 1) - pad[2], xh * up - yh * dn + pad[2] - (up - 1)]
        dx = run_emu(emu, g[k + '.r'], fd, fu, None, dn, up, pp, gain * up ** 2 / dn ** 2, slope, None, flip=True,
                     sign_mode=2, signs=so, sx=-(len(fu) - 1) + pad[0], sy=-(len(fu) - 1) + pad[2], tile=(16, 16))
        assert dx.shape == x.shape
        assert rel_err(dx, g[k + '.dx']) < 1e-5, name


def test_emulated_skip_and_scale(emu, golden_ops):
    g = golden_ops
    k = 'flrelu.u2d2'
    c = g[k + '.cfg']
    rng = np.random.default_rng(0)
    skip = rng.standard_normal(g[k + '.y'].shape).astype(np.float32)
    y = run_emu(emu, g[k + '.x'], g[k + '.fu'], g[k + '.fd'], g[k + '.b'], 2, 2, [9, 8, 9, 8], float(c[6]), float(c[7]),
                256.0, skip=skip, out_scale=0.25)
    assert rel_err(y, (g[k + '.y'] + skip) * 0.25) < 3e-6
