"""Timeline of CTA 0 of the tcgen05 convolution (afcm_conv_tc_trace): per role the recorded (event, clock) pairs of a few output
tiles in the steady state -- where the TMA producer, the MMA issuer and the first epilogue group wait and for how long.

    python tools/conv_tc_trace.py [--cin 64] [--cout 64] [--size 276] [--batch 64] [--direct 0]
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from afcm_b200 import _lib  # noqa: E402
from afcm_b200.torch_utils.ops import conv2d_gradfix  # noqa: E402

NAMES = {1: 'tma: slot free', 10: 'mma: tile start', 11: 'mma: accumulator free', 12: 'mma: stage landed', 13: 'mma: stage issued',
         20: 'epi: waiting for tile', 21: 'epi: accumulator ready', 22: 'epi: stores issued', 23: 'epi: coefficients visible', 24: 'epi: accumulator chunk in registers'}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--cin', type=int, default=64)
    ap.add_argument('--cout', type=int, default=64)
    ap.add_argument('--size', type=int, default=276)
    ap.add_argument('--batch', type=int, default=64)
    ap.add_argument('--direct', type=int, default=0)
    ap.add_argument('--issuers', type=int, default=2)
    ap.add_argument('--from-tile', type=int, default=40)
    ap.add_argument('--tiles', type=int, default=3)
    args = ap.parse_args()
    dev = torch.device('cuda:0')
    conv2d_gradfix.set_conv_impl('tc', torch.float16, act=torch.float16)
    conv2d_gradfix.direct_nchw = bool(args.direct)
    _lib.lib().afcm_conv_tc_set_issuers(args.issuers)
    x = torch.randn(args.batch, args.cin, args.size, args.size, device=dev).half()
    w = torch.randn(args.cout, args.cin, 3, 3, device=dev)
    run = lambda: conv2d_gradfix.conv2d_native(x, w, 2, impl='tc', out_dtype=torch.float16)
    run(); torch.cuda.synchronize()
    buf = torch.zeros(3 * 4096, dtype=torch.int64, device=dev)
    _lib.lib().afcm_conv_tc_trace(buf.data_ptr())
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); run(); b.record(); torch.cuda.synchronize()
    _lib.lib().afcm_conv_tc_trace(None)
    print('time %.3f ms (pack + GEMM when not direct)' % a.elapsed_time(b))
    t = buf.cpu().numpy().reshape(3, 4096)
    roles = ['TMA', 'MMA', 'EPI']
    ev = sorted((int(v) & ((1 << 48) - 1), r, int(v) >> 48) for r in range(3) for v in t[r] if v)
    starts = [c for c, r, code in ev if r == 1 and code == 10]
    if len(starts) <= args.from_tile + args.tiles:
        args.from_tile = max(0, len(starts) - args.tiles - 1)
    t0, t1 = starts[args.from_tile], starts[args.from_tile + args.tiles]
    print('tiles recorded: %d; mean clocks per tile: %.0f' % (len(starts), (starts[-1] - starts[0]) / max(1, len(starts) - 1)))
    last = {r: None for r in range(3)}
    for c, r, code in ev:
        if t0 <= c <= t1:
            print('%8d  %s  %-28s %s' % (c - t0, roles[r], NAMES.get(code, str(code)), '' if last[r] is None else '(+%d)' % (c - last[r])))
        last[r] = c


if __name__ == '__main__':
    main()
