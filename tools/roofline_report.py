#!/usr/bin/env python
"""Per-layer roofline table of the forward path from the committed layer benches (batch 64, fp16 planes, one B200):

    python tools/roofline_report.py [r02] > profiles/r02_roofline_table.md

Inputs: profiles/<round>_layer_bench_flr_tc_fp16.json (tools/layer_bench.py --ops flrelu_tc,f16in,f16out,nobias),
profiles/<round>_layer_bench_conv_fp16.json (--ops conv_tc,conv_nchw,f16in,f16out), MEASURED_PEAKS.json (driver-written) or the
fallbacks bench.py uses.  The kernels are timed ALONE, so the tensor-core column is quoted against the BURST cuBLAS peak
(bf16_tflops); inside the sustained, power-capped step the same kernels run against the sustained peak (bench.py).  Algorithmic work per SURVEY.md 8(d): filtered_lrelu bytes = 2 B x (Hc^2 + out^2) per plane,
pack bytes = 2 B read + 2 B written per (pixel, channel), convolution FLOP = 2 x Co x Ci x 9 x (H+2)^2 per slice."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    import sys
    rnd = sys.argv[1] if len(sys.argv) > 1 else 'r02'
    pk = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    peaks = json.load(open(pk)) if os.path.exists(pk) else {}
    hbm = float(peaks.get('hbm_gbs', 6650.0))
    tf = float(peaks.get('bf16_tflops', 1590.0))                 # isolated launches: burst peak
    flr = json.load(open(os.path.join(ROOT, 'profiles', rnd + '_layer_bench_flr_tc_fp16.json')))
    conv = json.load(open(os.path.join(ROOT, 'profiles', rnd + '_layer_bench_conv_fp16.json')))
    B = flr['summary']['batch']
    assert conv['summary']['batch'] == B
    crow = {r['layer']: r for r in conv['rows']}
    print(f'# Forward path, per layer (batch {B}, fp16 planes, one B200; isolated launches, L2 flushed between iterations)\n')
    print(f'Peaks: HBM copy {hbm:.0f} GB/s, dense bf16 {tf:.0f} TFLOP/s (burst: the kernels are timed alone) -- '
          f'{"MEASURED_PEAKS.json" if peaks else "fallback values"}.\n')
    print('| layer | conv Ci→Co @H | packed GEMM ms | TFLOP/s | of tensor peak | pack ms | GB/s | of HBM | direct GEMM ms (no pack pass) | TFLOP/s | of tensor peak | filtered_lrelu Hc→out (up/down) | ms | GB/s | of HBM |')
    print('|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|')
    tot = dict(conv=0.0, pack=0.0, flr=0.0, flops=0.0, pbytes=0.0, fbytes=0.0, direct=0.0)
    for r in flr['rows']:
        c = crow.get(r['layer'], {})
        cells = [r['layer'], f"{r['cin']}→{r['cout']} @{r['H']}"]
        if 'conv_tc_ms' in c:
            flops = 2.0 * B * c['cout'] * c['cin'] * 9 * c['Hc'] ** 2
            pbytes = c['pack_gbs'] * c['pack_ms'] * 1e6
            cells += [f"{c['conv_tc_ms']:.3f}", f"{c['conv_tc_tflops']:.0f}", f"{c['conv_tc_tflops'] / tf:.2f}",
                      f"{c['pack_ms']:.3f}", f"{c['pack_gbs']:.0f}", f"{c['pack_gbs'] / hbm:.2f}"]
            tot['conv'] += c['conv_tc_ms']; tot['pack'] += c['pack_ms']; tot['flops'] += flops; tot['pbytes'] += pbytes
            d = c.get('conv_pitched_ms', c.get('conv_nchw_ms'))
            if d:
                cells += [f"{d:.3f}", f"{flops / d / 1e9:.0f}", f"{flops / d / 1e9 / tf:.2f}"]
                tot['direct'] += d
            else:
                cells += ['-'] * 3
        else:
            cells += ['-'] * 9
        if 'flrelu_tc_ms' in r:
            fb = r['flrelu_tc_gbs'] * r['flrelu_tc_ms'] * 1e6
            cells += [f"{r['Hc']}→{r['out']} ({r['up']}/{r['down']})", f"{r['flrelu_tc_ms']:.3f}", f"{r['flrelu_tc_gbs']:.0f}",
                      f"{r['flrelu_tc_gbs'] / hbm:.2f}"]
            tot['flr'] += r['flrelu_tc_ms']; tot['fbytes'] += fb
        else:
            cells += ['-'] * 4
        print('| ' + ' | '.join(cells) + ' |')
    print(f"| **total** | | **{tot['conv']:.2f}** | **{tot['flops'] / tot['conv'] / 1e9:.0f}** | **{tot['flops'] / tot['conv'] / 1e9 / tf:.2f}** | "
          f"**{tot['pack']:.2f}** | **{tot['pbytes'] / tot['pack'] / 1e6:.0f}** | **{tot['pbytes'] / tot['pack'] / 1e6 / hbm:.2f}** | "
          f"**{tot['direct']:.2f}** | **{tot['flops'] / max(tot['direct'], 1e-9) / 1e9:.0f}** | **{tot['flops'] / max(tot['direct'], 1e-9) / 1e9 / tf:.2f}** | | "
          f"**{tot['flr']:.2f}** | **{tot['fbytes'] / tot['flr'] / 1e6:.0f}** | **{tot['fbytes'] / tot['flr'] / 1e6 / hbm:.2f}** |")
    print(f"\nSum per batch of {B}: packed path (pack + GEMM + filtered_lrelu) {tot['conv'] + tot['pack'] + tot['flr']:.1f} ms, direct path (GEMM reading "
          f"the NCHW planes + filtered_lrelu) {tot['direct'] + tot['flr']:.1f} ms; the measured step (profiles/{rnd}_bench_n1.json) adds the small kernels "
          f"and runs warm and power-capped.")


if __name__ == '__main__':
    main()
