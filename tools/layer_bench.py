"""Per-layer timing of the two heavy operators at the AFCM geometries -- a development
tool, not the bench contract.  CUDA-event timing on the launching stream, L2 flushed between iterations.

    python tools/layer_bench.py --batch 8 --ops flrelu,conv_tc,conv_f32
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from afcm_b200 import _lib  # noqa: E402
from afcm_b200.networks_stylegan3 import afcm_generator  # noqa: E402
from afcm_b200.torch_utils.ops import conv2d_gradfix  # noqa: E402
from afcm_b200.torch_utils.ops.filtered_lrelu import _run_fused, filtered_lrelu_tc  # noqa: E402


ITERS, WARMUP = 5, 2


def time_cuda(fn, iters=None, warmup=None, flush=None):
    iters = ITERS if iters is None else iters
    warmup = WARMUP if warmup is None else warmup
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=8)
    ap.add_argument('--ops', default='flrelu,conv_tc')
    ap.add_argument('--json', default='')
    ap.add_argument('--tile', default='')
    ap.add_argument('--layers', default='', help='comma-separated layer names (prefix match on enc<i> / L<i>_); default all')
    ap.add_argument('--iters', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=2)
    args = ap.parse_args()
    ops = args.ops.split(',')
    global ITERS, WARMUP
    ITERS, WARMUP = args.iters, args.warmup
    dev = torch.device('cuda:0')
    peaks = {}
    pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    hbm = peaks.get('hbm_gbs', 6650.0); tf = peaks.get('bf16_tflops', 1590.0)
    G = afcm_generator(seed=0, device=dev)
    S = G.synthesis
    layers = [('enc%d' % i, getattr(S, 'encoder_%d' % i)) for i in range(S.num_layers)] + \
             [(n, getattr(S, n)) for n in S.layer_names]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    B = args.batch
    if os.environ.get('AFCM_TC_ROWREUSE'):
        _lib.lib().afcm_conv_tc_set_rowreuse(int(os.environ['AFCM_TC_ROWREUSE']))
    if os.environ.get('AFCM_TC_ISSUERS'):
        _lib.lib().afcm_conv_tc_set_issuers(int(os.environ['AFCM_TC_ISSUERS']))
    if os.environ.get('AFCM_TC_DBG'):
        _lib.lib().afcm_conv_tc_debug_buffer(int(os.environ['AFCM_TC_DBG']) << 8)
    if os.environ.get('AFCM_FTC_WAVES'):
        _lib.lib().afcm_filtered_lrelu_tc_set_waves(int(os.environ['AFCM_FTC_WAVES']))
    if args.tile:
        tw, th = [int(v) for v in args.tile.split('x')]
        _lib.lib().afcm_filtered_lrelu_set_tile(tw, th)
    rows = []
    tot = dict(flrelu_ms=0.0, flrelu_bytes=0.0, conv_tc_ms=0.0, pack_ms=0.0, conv_f32_ms=0.0, flops=0.0)
    want = [w for w in args.layers.split(',') if w]
    for name, L in layers:
        if want and not any(name == w or name.startswith(w + '_') for w in want):
            continue
        cin, cout = L.in_channels, L.out_channels
        H = int(L.in_size[0]); k = L.conv_kernel; Hc = H + k - 1; out = int(L.out_size[0])
        row = dict(layer=name, cin=cin, cout=cout, H=H, Hc=Hc, out=out, up=L.up_factor, down=L.down_factor)
        if 'flrelu' in ops:
            x = torch.randn(B, cout, Hc, Hc, device=dev)
            b = torch.randn(cout, device=dev)
            px0, px1, py0, py1 = L.padding
            gain, slope = (1.0, 1.0) if getattr(L, 'is_torgb', False) else (float(np.sqrt(2)), 0.2)
            fn = lambda: _run_fused(x, L.up_filter, L.down_filter, b, None, L.up_factor, L.down_factor, px0, px1, py0, py1,
                                    0, 0, gain, slope, 256.0, False, False)
            ms = time_cuda(fn, flush=flush)
            nbytes = 4.0 * B * cout * (Hc * Hc + out * out)
            row.update(flrelu_ms=ms, flrelu_gbs=nbytes / ms / 1e6, flrelu_frac=nbytes / ms / 1e6 / hbm)
            tot['flrelu_ms'] += ms; tot['flrelu_bytes'] += nbytes
            del x
        if 'flrelu_tc' in ops and k == 3:
            x = torch.randn(B, cout, Hc, Hc, device=dev)
            b = torch.randn(cout, device=dev)
            out_dt = torch.float16 if 'f16out' in ops else torch.float32
            if 'f16in' in ops:
                x = x.half()
            if 'nobias' in ops:
                b = None
            impl = 'tc'
            if 't5' in ops:                       # the tcgen05 / TMEM kernel: fp16 planes with a 16-byte aligned row pitch, no bias
                from afcm_b200.torch_utils.ops.filtered_lrelu import padded_pitch_empty
                xv = padded_pitch_empty(x.shape, torch.float16, dev)
                xv.copy_(x)
                x, b, impl = xv, None, 't5'
            fn = lambda: filtered_lrelu_tc(x, L.up_filter, L.down_filter, b, up=L.up_factor, down=L.down_factor,
                                           padding=L.padding, gain=float(np.sqrt(2)), slope=0.2, clamp=256.0, out_dtype=out_dt, impl=impl)
            assert fn() is not None
            ms = time_cuda(fn, flush=flush)
            nbytes = float(B * cout * (x.element_size() * Hc * Hc + (2 if out_dt == torch.float16 else 4) * out * out))
            row.update(flrelu_tc_ms=ms, flrelu_tc_gbs=nbytes / ms / 1e6, flrelu_tc_frac=nbytes / ms / 1e6 / hbm)
            tot['flrelu_tc_ms'] = tot.get('flrelu_tc_ms', 0.0) + ms; tot['flrelu_tc_bytes'] = tot.get('flrelu_tc_bytes', 0.0) + nbytes
            del x
        flops = 2.0 * B * cout * cin * k * k * Hc * Hc
        tot['flops'] += flops
        if k == 3 and ('conv_tc' in ops or 'conv_f32' in ops):
            x = torch.randn(B, cin, H, H, device=dev)
            if 'f16in' in ops:
                x = x.half()
            w = L.weight.detach()
            if 'conv_tc' in ops:
                Lb = _lib.lib()
                ent = conv2d_gradfix.prepare_weight(w, 1.0, False, want_tc=True)
                plane = int(Lb.afcm_conv_tc_plane_elems(H, H, cin))
                xp = torch.empty(B, plane, dtype=torch.float16, device=dev)
                y = torch.empty(B, cout, Hc, Hc, device=dev, dtype=torch.float16 if 'f16out' in ops else torch.float32)
                st = _lib.stream_ptr(dev)
                pack = lambda: _lib.check(Lb.afcm_conv_tc_pack(_lib.ptr(x), _lib.dtype_code(x.dtype), None, _lib.ptr(xp), 1, B, cin, H, H, st))
                gemm = lambda: _lib.check(Lb.afcm_conv2d_tc(_lib.ptr(xp), _lib.ptr(ent[('w_tc', torch.float16)]), None, None,
                                                            _lib.ptr(y), _lib.dtype_code(y.dtype), 1, B, cin, H, H, cout, 2, st))
                pms = time_cuda(pack, flush=flush); gms = time_cuda(gemm, flush=flush)
                pbytes = float(x.element_size() * x.numel() + 2 * xp.numel())
                row.update(pack_ms=pms, pack_gbs=pbytes / pms / 1e6, conv_tc_ms=gms, conv_tc_tflops=flops / gms / 1e9, conv_tc_frac=flops / gms / 1e9 / tf)
                tot['conv_tc_ms'] += gms; tot['pack_ms'] += pms
                if 'conv_nchw' in ops and x.dtype == torch.float16:
                    # the GEMM reading the fp16 NCHW planes itself (no pack pass)
                    direct = lambda: _lib.check(Lb.afcm_conv2d_tc_nchw(_lib.ptr(x), H, None, _lib.ptr(ent[('w_tc', torch.float16)]), None, None,
                                                                       _lib.ptr(y), _lib.dtype_code(y.dtype), B, cin, H, H, cout, st))
                    dms = time_cuda(direct, flush=flush)
                    row.update(conv_nchw_ms=dms, conv_nchw_tflops=flops / dms / 1e9, conv_nchw_frac=flops / dms / 1e9 / tf)
                    tot['conv_nchw_ms'] = tot.get('conv_nchw_ms', 0.0) + dms
                    # planes stored at the pitch W + 2 with zero pad columns (what filtered_lrelu_tc writes on the fast path)
                    xq = torch.zeros(B, cin, H, H + 2, device=dev, dtype=torch.float16)
                    xq[..., :H] = x
                    pitched = lambda: _lib.check(Lb.afcm_conv2d_tc_nchw(_lib.ptr(xq), H + 2, None, _lib.ptr(ent[('w_tc', torch.float16)]), None, None,
                                                                        _lib.ptr(y), _lib.dtype_code(y.dtype), B, cin, H, H, cout, st))
                    qms = time_cuda(pitched, flush=flush)
                    row.update(conv_pitched_ms=qms)
                    tot['conv_pitched_ms'] = tot.get('conv_pitched_ms', 0.0) + qms
                    del xq
                del xp, y
            if 'conv_f32' in ops:
                fms = time_cuda(lambda: conv2d_gradfix.conv2d_native(x.float(), w, 2, impl='f32'), iters=2, warmup=1)
                row.update(conv_f32_ms=fms, conv_f32_tflops=flops / fms / 1e9)
                tot['conv_f32_ms'] += fms
            del x
        rows.append(row)
        print(json.dumps(row))
    summ = dict(batch=B, **tot)
    if tot['flrelu_ms']:
        summ['flrelu_gbs'] = tot['flrelu_bytes'] / tot['flrelu_ms'] / 1e6
        summ['flrelu_frac_of_measured_hbm'] = summ['flrelu_gbs'] / hbm
        summ['flrelu_ms_per_slice'] = tot['flrelu_ms'] / B
    if tot.get('flrelu_tc_ms'):
        summ['flrelu_tc_gbs'] = tot['flrelu_tc_bytes'] / tot['flrelu_tc_ms'] / 1e6
        summ['flrelu_tc_frac_of_measured_hbm'] = summ['flrelu_tc_gbs'] / hbm
        summ['flrelu_tc_ms_per_slice'] = tot['flrelu_tc_ms'] / B
    if tot.get('conv_nchw_ms'):
        summ['conv_nchw_tflops'] = tot['flops'] / tot['conv_nchw_ms'] / 1e9
    if tot['conv_tc_ms']:
        summ['conv_tc_tflops'] = tot['flops'] / tot['conv_tc_ms'] / 1e9
        summ['conv_tc_frac_of_measured_bf16'] = summ['conv_tc_tflops'] / tf
        summ['conv_tc_ms_per_slice'] = tot['conv_tc_ms'] / B
    print('SUMMARY', json.dumps(summ))
    if args.json:
        json.dump(dict(rows=rows, summary=summ), open(args.json, 'w'), indent=1)


if __name__ == '__main__':
    main()
