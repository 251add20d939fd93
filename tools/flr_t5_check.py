"""Development check of the tcgen05 / TMEM filtered_lrelu (afcm_filtered_lrelu_t5) on a GPU: every case is compared with the
CPU oracle, errors are summarised per case (max error, where, row / column error profiles) and the run continues after a
numerical failure, so that one GPU call tells as much as possible.

    python tools/flr_t5_check.py [--big] [--bench BATCH]
"""
import argparse
import os
import sys
import time

import numpy as np
import scipy.signal
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from afcm_b200.torch_utils.ops import filtered_lrelu as flr  # noqa: E402
from oracle import afcm_oracle as orc  # noqa: E402

CASES = [  # up, down, padding, N, C, H, W
    (2, 2, [9, 8, 9, 8], 1, 1, 22, 26),
    (2, 2, [9, 8, 9, 8], 2, 3, 38, 38),
    (2, 2, [9, 8, 9, 8], 1, 2, 70, 150),
    (2, 2, [9, 8, 9, 8], 1, 2, 150, 150),
    (2, 2, [-11, -12, -11, -12], 1, 3, 38, 36),
    (2, 2, [9, 8, 7, 10], 2, 2, 21, 20),
    (2, 2, [8, 9, 10, 7], 1, 2, 70, 84),
    (4, 2, [-6, -9, -6, -9], 2, 3, 22, 26),
    (4, 2, [-6, -9, -6, -9], 1, 4, 38, 38),
    (4, 2, [3, 2, 1, 4], 1, 3, 9, 12),
    (4, 2, [-5, -10, -7, -8], 1, 2, 54, 54),
    (4, 2, [-6, -9, -6, -9], 1, 2, 86, 86),
    (2, 4, [34, 33, 34, 33], 2, 3, 38, 42),
    (2, 4, [34, 33, 34, 33], 1, 2, 54, 54),
    (2, 4, [33, 34, 35, 32], 1, 2, 86, 86),
    (2, 4, [34, 33, 34, 33], 1, 2, 150, 150),
]
BIG = [
    (2, 2, [9, 8, 9, 8], 1, 3, 278, 278),
    (2, 4, [34, 33, 34, 33], 1, 3, 278, 278),
    (4, 2, [-6, -9, -6, -9], 1, 3, 150, 150),
    (2, 2, [-11, -12, -11, -12], 1, 3, 278, 278),
]


def filters(up, down):
    fu = scipy.signal.firwin(6 * up, 0.4, width=0.3, fs=2).astype(np.float32)
    fd = scipy.signal.firwin(6 * down, 0.25, width=0.2, fs=2).astype(np.float32)
    return fu, fd


def padded(x16, dev):
    N, C, H, W = x16.shape
    v = flr.padded_pitch_empty([N, C, H, W], torch.float16, dev)
    v.copy_(torch.from_numpy(x16).to(dev))
    return v


def run_case(case, dev, clamp=3.0, skip=False):
    up, down, pad, N, C, H, W = case
    rng = np.random.RandomState(H * 11 + W)
    fu, fd = filters(up, down)
    x = (rng.randn(N, C, H, W) * 2).astype(np.float16)
    ref = orc.filtered_lrelu(x.astype(np.float32), fu, fd, None, up=up, down=down, padding=pad, gain=np.sqrt(2), slope=0.2, clamp=clamp)
    ref = ref[0] if isinstance(ref, tuple) else ref
    sk = None
    scale = 1.0
    if skip:
        sk = rng.randn(*ref.shape).astype(np.float16)
        scale = 0.25
        ref = (ref + sk.astype(np.float32)) * scale
    xv = padded(x, dev)
    t = lambda a: torch.from_numpy(a).to(dev)
    y = flr.filtered_lrelu_tc(xv, t(fu), t(fd), None, up=up, down=down, padding=pad, gain=np.sqrt(2), slope=0.2, clamp=clamp,
                              out_dtype=torch.float16, skip=None if sk is None else t(sk), out_scale=scale, impl='t5')
    torch.cuda.synchronize()
    if y is None:
        print('  UNSUPPORTED:', flr._lib.last_error())
        return False
    y = y.float().cpu().numpy()
    d = np.abs(y - ref)
    err = d.max() / np.abs(ref).max()
    ok = err <= 4e-3
    print(f'case up={up} down={down} pad={pad} N={N} C={C} {H}x{W} -> {ref.shape[2]}x{ref.shape[3]} skip={skip}: rel max err {err:.3e} {"ok" if ok else "FAIL"}')
    if not ok:
        idx = np.unravel_index(np.argmax(d), d.shape)
        print('   worst at', idx, 'got', y[idx], 'want', ref[idx])
        bad = d > 4e-3 * np.abs(ref).max()
        rows = np.where(bad.any(axis=(0, 1, 3)))[0]; cols = np.where(bad.any(axis=(0, 1, 2)))[0]
        print('   bad rows:', rows[:40], '... count', len(rows), ' bad cols:', cols[:40], '... count', len(cols))
        print('   bad planes:', np.where(bad.any(axis=(2, 3)))[0][:8], np.where(bad.any(axis=(2, 3)))[1][:8], 'nan count', int(np.isnan(y).sum()))
    return ok


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--big', action='store_true')
    ap.add_argument('--only', type=int, default=-1)
    args = ap.parse_args()
    dev = torch.device('cuda:0')
    cases = CASES + (BIG if args.big else [])
    if args.only >= 0:
        cases = [cases[args.only]]
    nfail = 0
    for i, c in enumerate(cases):
        try:
            ok = run_case(c, dev)
            if i in (1, 8, 13):
                ok = run_case(c, dev, clamp=256.0, skip=True) and ok
        except RuntimeError as e:
            print('CUDA / library error on case', c, ':', str(e)[:300])
            nfail += 1
            break
        nfail += 0 if ok else 1
    print('FAILED' if nfail else 'ALL OK', nfail)
    return 1 if nfail else 0


if __name__ == '__main__':
    sys.exit(main())
