"""Timeline of CTA 0 of the tcgen05 / TMEM filtered_lrelu (afcm_filtered_lrelu_t5_trace): runs one layer geometry and prints,
per warp role, the recorded (event, clock) pairs of a few steps in the steady state -- where each role waits and for how long.

    python tools/flr_t5_trace.py [--size 278] [--up 2] [--down 2] [--planes 2048] [--from-step 20] [--steps 3]
"""
import argparse
import os
import sys

import numpy as np
import scipy.signal
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from afcm_b200 import _lib  # noqa: E402
from afcm_b200.torch_utils.ops import filtered_lrelu as flr  # noqa: E402

NAMES = {1: 'tma: slot free', 2: 'tma: issued', 10: 'mma: input landed + D1 free', 11: 'mma: P1 issued', 12: 'mma: P1 complete',
         40: 'mma: P4 may start', 41: 'mma: P4 issued', 60: 'wg: D3 complete', 61: 'wg: E2 done', 70: 'wg: D4 complete', 71: 'wg: E3 done'}
for g in range(5):
    NAMES[20 + g] = f'mma: D2 buffer free (group {g})'; NAMES[25 + g] = f'mma: P2 issued (group {g})'
    NAMES[30 + g] = f'mma: A3 ready (group {g})'; NAMES[35 + g] = f'mma: P3 issued (group {g})'
    NAMES[50 + g] = f'wg: D2 ready (group {g})'; NAMES[55 + g] = f'wg: A3 written (group {g})'


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--size', type=int, default=278)
    ap.add_argument('--up', type=int, default=2)
    ap.add_argument('--down', type=int, default=2)
    ap.add_argument('--planes', type=int, default=2048)
    ap.add_argument('--from-step', type=int, default=20)
    ap.add_argument('--steps', type=int, default=2)
    args = ap.parse_args()
    dev = torch.device('cuda:0')
    pad = {(2, 2): [9, 8, 9, 8], (4, 2): [-6, -9, -6, -9], (2, 4): [34, 33, 34, 33]}[(args.up, args.down)]
    fu = torch.from_numpy(scipy.signal.firwin(6 * args.up, 0.4, width=0.3, fs=2).astype(np.float32)).to(dev)
    fd = torch.from_numpy(scipy.signal.firwin(6 * args.down, 0.25, width=0.2, fs=2).astype(np.float32)).to(dev)
    x = flr.padded_pitch_empty([1, args.planes, args.size, args.size], torch.float16, dev)
    x.normal_()
    buf = torch.zeros(4 * 2048, dtype=torch.int64, device=dev)      # one word per event: code << 48 | clock
    run = lambda: flr.filtered_lrelu_tc(x, fu, fd, None, up=args.up, down=args.down, padding=pad, clamp=256.0, out_dtype=torch.float16, impl='t5')
    run(); torch.cuda.synchronize()
    _lib.lib().afcm_filtered_lrelu_t5_trace(buf.data_ptr())
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); y = run(); b.record(); torch.cuda.synchronize()
    _lib.lib().afcm_filtered_lrelu_t5_trace(None)
    print('kernel time %.3f ms, output %s' % (a.elapsed_time(b), tuple(y.shape)))
    t = buf.cpu().numpy().reshape(4, 2048)
    roles = ['TMA', 'MMA', 'WG0', 'WG1']
    ev = []
    for r in range(4):
        for v in t[r]:
            if v:
                ev.append((int(v) & ((1 << 48) - 1), r, int(v) >> 48))
    ev.sort()
    # step boundaries = code 10 events of the MMA role
    p1 = [c for c, r, code in ev if r == 1 and code == 10]
    if len(p1) <= args.from_step + args.steps:
        args.from_step = max(0, len(p1) - args.steps - 1)
    t0, t1 = p1[args.from_step], p1[args.from_step + args.steps]
    print('steps recorded: %d; mean clocks per step over the record: %.0f' % (len(p1), (p1[-1] - p1[0]) / max(1, len(p1) - 1)))
    last = {r: None for r in range(4)}
    for c, r, code in ev:
        if t0 <= c <= t1:
            d = '' if last[r] is None else '(+%d)' % (c - last[r])
            print('%8d  %s  %-40s %s' % (c - t0, roles[r], NAMES.get(code, str(code)), d))
        last[r] = c


if __name__ == '__main__':
    main()
