#!/usr/bin/env python
"""The kernel to beat: the REFERENCE's own GPU path timed on this box's B200 -- its JIT-built CUDA plugins
(filtered_lrelu_plugin, bias_act_plugin, upfirdn2d_plugin through torch_utils/custom_ops.get_plugin, custom_ops.py:59-155) and
cuDNN for the (grouped) convolutions, fp32 tensors, exactly as `Stylegan3Generator.forward` runs them.  Reported next to this
library's numbers; nothing here is part of the product path.

    python tools/ref_cuda_bench.py [--batch 16] [--steps 3] [--json out.json]

Needs the staged reference tree (tools/stage_reference.py), nvcc and ninja (the plugins compile at first use, a few minutes).
Prints per-operator timings of filtered_lrelu at the AFCM geometries and the whole-generator throughput; says so when a plugin
does not build (the reference then silently takes its slow `_ref` composition, which is still what a user of the reference on
this box would get)."""
import argparse
import importlib
import json
import os
import sys
import time
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIRS = [os.path.join(ROOT, 'baseline', '_ref', 'AFCM'), '/root/reference']


def cuda_ms(fn, iters=3, warmup=1):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=16)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--json', default='')
    args = ap.parse_args()
    ref = next((d for d in REF_DIRS if os.path.isdir(os.path.join(d, 'models', 'networks', 'stylegan3'))), None)
    if ref is None:
        print(json.dumps(dict(unavailable='reference tree not staged (python tools/stage_reference.py)')))
        return 0
    sys.dont_write_bytecode = True
    sys.path.insert(0, ref)
    os.environ.setdefault('TORCH_CUDA_ARCH_LIST', '10.0')
    warnings.filterwarnings('ignore')
    net = importlib.import_module('models.networks.stylegan3.networks_stylegan3')
    ops = importlib.import_module('models.networks.stylegan3.torch_utils.ops.filtered_lrelu')
    custom_ops = importlib.import_module('models.networks.stylegan3.torch_utils.custom_ops')
    custom_ops.verbosity = 'brief'
    # Environment shim, not a change of the reference: its loader (custom_ops.py:136-142, written for torch 1.9) builds with
    # torch.utils.cpp_extension.load() and then import_module()s the plugin by name; torch 2.x no longer leaves the built module
    # importable by name, so the module load() returns is registered under that name.
    import torch.utils.cpp_extension as cpp_ext
    _load = cpp_ext.load

    def load_and_register(name, *a, **kw):
        mod = _load(name, *a, **kw)
        sys.modules[name] = mod
        return mod
    cpp_ext.load = load_and_register
    dev = torch.device('cuda:0')
    out = dict(device=torch.cuda.get_device_name(0), torch=torch.__version__, batch=args.batch,
               allow_tf32_conv=bool(torch.backends.cudnn.allow_tf32))
    t0 = time.time()
    try:
        plugin_ok = bool(ops._init())
    except Exception as e:
        plugin_ok = False
        out['plugin_error'] = repr(e)[:300]
    out['filtered_lrelu_plugin_built'] = plugin_ok
    out['plugin_build_s'] = time.time() - t0

    torch.manual_seed(0)
    G = net.Stylegan3Generator(
        z_dim=512, c_dim=1, w_dim=512, img_resolution=256, img_channels_in=4, img_channels_out=1, mapping_kwargs=dict(num_layers=8),
        synthesis_kwargs=dict(channel_base=16384, channel_max=512, num_layers=14, num_critical=2, first_cutoff=2,
                              first_stopband=2 ** 2.1, last_stopband_rel=2 ** 0.3, margin_size=10, output_scale=0.25,
                              skip_resolution=128, conv_kernel=3, filter_size=6, lrelu_upsampling=2, use_radial_filters=False,
                              conv_clamp=256, magnitude_ema_beta=0.5 ** (16 / (20 * 1e3)), cond_mod=True)).eval().to(dev)
    S = G.synthesis
    B = args.batch
    # ---- filtered_lrelu per layer geometry, the reference's CUDA op (or its fallback), fp32
    layers = [('enc%d' % i, getattr(S, 'encoder_%d' % i)) for i in range(S.num_layers)] + [(n, getattr(S, n)) for n in S.layer_names]
    rows, tot_ms, tot_bytes = [], 0.0, 0.0
    for name, L in layers:
        if getattr(L, 'is_torgb', False):
            continue
        C, Hc, o = L.out_channels, int(L.in_size[0]) + 2, int(L.out_size[0])
        x = torch.randn(B, C, Hc, Hc, device=dev)
        b = torch.randn(C, device=dev)
        fn = lambda: ops.filtered_lrelu(x=x, fu=L.up_filter, fd=L.down_filter, b=b, up=L.up_factor, down=L.down_factor, padding=L.padding,
                                        gain=np.sqrt(2), slope=0.2, clamp=256)
        with torch.no_grad():
            ms = cuda_ms(fn)
        nbytes = 4.0 * B * C * (Hc * Hc + o * o)
        rows.append(dict(layer=name, C=C, Hc=Hc, out=o, up=L.up_factor, down=L.down_factor, ms=ms, gbs=nbytes / ms / 1e6))
        tot_ms += ms; tot_bytes += nbytes
        del x
    out['filtered_lrelu'] = dict(rows=rows, ms_per_batch=tot_ms, ms_per_slice=tot_ms / B, gbs=tot_bytes / tot_ms / 1e6,
                                 note='reference CUDA plugin (fp32 I/O, 925.9 MB/slice algorithmic)' if plugin_ok else 'plugin did not build: _ref composition')
    # ---- the whole generator forward through the reference's stock path
    g = torch.Generator().manual_seed(1)
    z = torch.randn(B, 512, generator=g).to(dev); c = torch.zeros(B, 1, device=dev)
    x = (torch.rand(B, 4, 256, 256, generator=g) * 2 - 1).to(dev)
    with torch.no_grad():
        ms = cuda_ms(lambda: G(z, c, x, noise_mode='const'), iters=args.steps)
    out['generator'] = dict(ms_per_step=ms, slices_per_sec=B / (ms * 1e-3), batch=B,
                            note='Stylegan3Generator.forward, fp32, reference plugins + cuDNN (TF32 convolutions allowed by torch default)')
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    with torch.no_grad():
        ms = cuda_ms(lambda: G(z, c, x, noise_mode='const'), iters=args.steps)
    out['generator_fp32_strict'] = dict(ms_per_step=ms, slices_per_sec=B / (ms * 1e-3), batch=B, note='the same with TF32 disabled')
    print(json.dumps(out))
    if args.json:
        json.dump(out, open(args.json, 'w'), indent=1)
    return 0


if __name__ == '__main__':
    sys.exit(main())
