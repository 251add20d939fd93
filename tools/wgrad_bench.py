#!/usr/bin/env python
"""Per-layer timing of the convolution gradients at the AFCM layer shapes (batch 32): weight gradient (mma.sync),
data gradient and forward (tcgen05).  python tools/wgrad_bench.py [--batch 32]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from afcm_b200 import _lib  # noqa: E402
from afcm_b200.torch_utils.ops import conv2d_gradfix as cg  # noqa: E402

# (Ci, Co, H) of the 29 3x3 convolutions, pad 2 except e_16x16 (pad 1)
LAYERS = [(4, 64, 276), (64, 64, 276), (64, 91, 276), (91, 128, 276), (128, 181, 276), (181, 256, 148), (256, 362, 148),
          (362, 512, 148), (512, 512, 84), (512, 512, 84), (512, 512, 52), (512, 512, 52), (512, 512, 36), (512, 512, 36),
          (512, 512, 36, 1), (512, 512, 36), (512, 512, 36), (512, 512, 36), (512, 512, 36), (512, 512, 52), (512, 512, 52),
          (512, 512, 84), (512, 362, 84), (362, 256, 148), (256, 181, 148), (181, 128, 148), (128, 91, 276), (91, 64, 276),
          (64, 64, 276)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=32)
    ap.add_argument('--reps', type=int, default=3)
    args = ap.parse_args()
    dev = torch.device('cuda:0')
    L = _lib.lib()
    dt = torch.bfloat16
    code = _lib.dtype_code(dt)
    st = _lib.stream_ptr(dev)
    N = args.batch
    rows, tot = [], dict(wgrad=0.0, wgrad_mma=0.0, dgrad=0.0, fwd=0.0, flops=0.0)
    seen = {}
    for ent in LAYERS:
        Ci, Co, H = ent[:3]
        pad = ent[3] if len(ent) > 3 else 2
        key = (Ci, Co, H, pad)
        if key not in seen:
            OH = H + 2 * pad - 2
            x = torch.randn(N, Ci, H, H, device=dev)
            dy = torch.randn(N, Co, OH, OH, device=dev)
            w = torch.randn(Co, Ci, 3, 3, device=dev) / (3 * Ci ** 0.5)
            xp, dyp = cg._pack(x, None, dt), cg._pack(dy, None, dt)
            w_tc = cg._weight_tc(w, dt)
            wT_tc = cg._weight_tc(w.flip(2, 3).transpose(0, 1).contiguous(), dt)
            dw = torch.empty_like(w)
            y = torch.empty(N, Co, OH, OH, device=dev)
            dx = torch.empty(N, Ci, H, H, device=dev)
            nbytes = int(L.afcm_conv2d_wgrad_tc_workspace(N, Ci, H, H, Co, pad))
            ws = torch.empty(nbytes // 4, device=dev)
            fns = dict(
                wgrad=lambda: _lib.check(L.afcm_conv2d_wgrad_tc5(_lib.ptr(dyp), _lib.ptr(xp), _lib.ptr(dw), _lib.ptr(ws), nbytes, code, N, Ci, H, H, Co, pad, st)),
                wgrad_mma=lambda: _lib.check(L.afcm_conv2d_wgrad_tc(_lib.ptr(dyp), _lib.ptr(xp), _lib.ptr(dw), code, N, Ci, H, H, Co, pad, st)),
                dgrad=lambda: _lib.check(L.afcm_conv2d_tc(_lib.ptr(dyp), _lib.ptr(wT_tc), None, None, _lib.ptr(dx), _lib.F32, code, N, Co, OH, OH, Ci, 2 - pad, st)),
                fwd=lambda: _lib.check(L.afcm_conv2d_tc(_lib.ptr(xp), _lib.ptr(w_tc), None, None, _lib.ptr(y), _lib.F32, code, N, Ci, H, H, Co, pad, st)))
            res = {}
            for name, fn in fns.items():
                fn()
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(args.reps):
                    fn()
                b.record()
                torch.cuda.synchronize()
                res[name] = a.elapsed_time(b) / args.reps
            flops = 2.0 * N * Co * Ci * 9 * OH * OH
            seen[key] = (res, flops)
            del x, dy, xp, dyp, y, dx
        res, flops = seen[key]
        rows.append(dict(Ci=Ci, Co=Co, H=H, pad=pad, gflop=flops / 1e9, **{k + '_ms': v for k, v in res.items()},
                         **{k + '_tflops': flops / v / 1e9 for k, v in res.items()}))
        for k in ('wgrad', 'wgrad_mma', 'dgrad', 'fwd'):
            tot[k] += res[k]
        tot['flops'] += flops
    for r in rows:
        print('%4d -> %4d @%3d pad %d  %7.1f GF   wgrad tcgen05 %6.2f ms %6.1f TF/s  mma.sync %6.2f ms %6.1f   dgrad %6.2f ms %6.1f   fwd %6.2f ms %6.1f' % (
            r['Ci'], r['Co'], r['H'], r['pad'], r['gflop'], r['wgrad_ms'], r['wgrad_tflops'], r['wgrad_mma_ms'], r['wgrad_mma_tflops'],
            r['dgrad_ms'], r['dgrad_tflops'], r['fwd_ms'], r['fwd_tflops']))
    print('total: wgrad tcgen05 %.1f ms (%.0f TF/s)  mma.sync %.1f ms (%.0f)  dgrad %.1f ms (%.0f)  fwd %.1f ms (%.0f)' % (
        tot['wgrad'], tot['flops'] / tot['wgrad'] / 1e9, tot['wgrad_mma'], tot['flops'] / tot['wgrad_mma'] / 1e9,
        tot['dgrad'], tot['flops'] / tot['dgrad'] / 1e9, tot['fwd'], tot['flops'] / tot['fwd'] / 1e9))
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    json.dump(dict(batch=N, rows=rows, total=tot), open(os.path.join(ROOT, 'gpurun_out', 'wgrad_bench.json'), 'w'), indent=1)


if __name__ == '__main__':
    main()
