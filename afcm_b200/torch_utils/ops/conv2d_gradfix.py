"""conv2d_gradfix -- the convolution entry point of the reference's operator API
(models/networks/stylegan3/torch_utils/ops/conv2d_gradfix.py:37-40, which forwards to cuDNN), served by
the hand-written kernels of libafcm_b200: an exact-fp32 SIMT kernel (`impl='f32'`, the parity path) and
the tcgen05/TMEM implicit GEMM with 16-bit operands and fp32 accumulation (`impl='tc'`).

Module switches (process-wide, like the reference's `enabled` flag):
    conv_impl : 'f32' | 'tc'      default implementation used by conv2d() / modulated_conv2d()
    tc_dtype  : torch.float16 | torch.bfloat16   operand type of the tensor-core path
    act_dtype : torch.float32 | torch.float16    storage type of the activations BETWEEN the operators of
                the generator (float16 = the fast inference path: tcgen05 conv -> fp16 NCHW -> tensor-core
                filtered_lrelu -> fp16 NCHW -> pack; only meaningful with conv_impl 'tc')
"""
import numpy as np
import os

import torch

from ... import _lib

conv_impl = 'f32'
tc_dtype = torch.float16
act_dtype = torch.float32
direct_nchw = True                   # tensor-core path, fp16 activations, full padding: the GEMM reads the NCHW planes itself (no pack pass)
hybrid_pack = os.environ.get('AFCM_HYBRID_PACK', '0') == '1'   # opt-in: pack + GEMM on the layers where it is faster in isolation, see _prefers_pack


def _prefers_pack(Ci, Co, H, W):
    """Direct-NCHW or pack + GEMM, per layer geometry (profiles/r02_layer_bench_conv_issuers1.json, batch 32): with few input
    channels on large planes the producer warps of the direct kernel, not the tensor core, set the pace (64 -> 64 channels at
    276 px: 0.51 ms direct against 0.13 + 0.29 ms), from ~128 input channels on the direct kernel wins or ties.  Inside the
    forward step the switch is neutral (60.1 against 60.4 ms at batch 64: -2.5 ms of GEMM, +1.8 ms of pack, and the step is
    power-capped), so it is off by default: one kernel per convolution."""
    return hybrid_pack and Ci <= 96 and H * W >= 200 * 200
wgrad_impl = 'tcgen05'               # tensor-core weight gradient: 'tcgen05' (TMEM accumulators) | 'mma' (mma.sync + atomics)
enabled = False                      # kept for API compatibility with the reference module
weight_gradients_disabled = False

_prep_cache = {}
_prep_generation = 0                 # bumped whenever a prepared tensor may have been released or replaced (see prep_generation())


def prep_generation():
    """Counter of events that invalidate device addresses handed out by prepare_weight(): cache clears / evictions and
    re-preparation after a weight update.  A captured CUDA graph has those addresses baked in; inference.GraphedGenerator
    records the value at capture time and re-captures when it has moved, so a replay never reads freed or stale weights."""
    return _prep_generation


def clear_prepared_weights():
    """Drops every prepared weight (after an optimizer step: the fp32 parameters have changed)."""
    global _prep_generation
    _prep_cache.clear()
    _prep_generation += 1


def set_conv_impl(impl, dtype=None, act=None):
    global conv_impl, tc_dtype, act_dtype
    assert impl in ('f32', 'tc')
    conv_impl = impl
    if dtype is not None:
        assert dtype in (torch.float16, torch.bfloat16)
        tc_dtype = dtype
    act = act or torch.float32
    assert act in (torch.float32, torch.float16)
    assert act == torch.float32 or impl == 'tc', 'fp16 activation storage needs the tensor-core path'
    act_dtype = act


def fast_path():
    """True when the generator runs its fast inference path (tcgen05 conv + tensor-core filtered_lrelu,
    fp16 activation storage)."""
    return conv_impl == 'tc' and act_dtype == torch.float16


def prepare_weight(w, pre_scale=1.0, normalize=False, want_tc=False, want_wsq=False, dtype=None):
    """Per-layer constant preparation (afcm_conv_weight_prep), cached on (storage, version): scaled /
    RMS-normalised fp32 weights, the 16-bit tap-major copy for the tensor-core kernel and the per-(o,i)
    squared sums used by demodulation."""
    import weakref
    dtype = dtype or tc_dtype
    key = (id(w), float(pre_scale), bool(normalize))
    ent = _prep_cache.get(key)
    global _prep_generation
    if ent is not None and (ent['_ref']() is not w or ent['_version'] != w._version):
        ent = None                      # address / id reuse or in-place update (optimizer step): re-prepare
        _prep_generation += 1
    if ent is None:
        if len(_prep_cache) > 512:
            _prep_generation += 1
            for k in [k for k, v in _prep_cache.items() if v['_ref']() is None]:
                del _prep_cache[k]
            if len(_prep_cache) > 512:
                _prep_cache.clear()
        ent = _prep_cache[key] = {'_ref': weakref.ref(w), '_version': w._version}
    need_f32 = 'w_f32' not in ent
    need_tc = want_tc and ('w_tc', dtype) not in ent
    need_wsq = want_wsq and 'wsq' not in ent
    if need_f32 or need_tc or need_wsq:
        Co, Ci, kh, kw = w.shape
        assert kh == kw
        wc = w.detach().contiguous().float()
        w_f32 = torch.empty_like(wc) if need_f32 else None
        w_tc = None
        if need_tc:
            w_tc = torch.empty([kh * kw, (Co + 15) // 16 * 16, (Ci + 63) // 64 * 64], dtype=dtype, device=w.device)
        wsq = torch.empty([Co, Ci], dtype=torch.float32, device=w.device) if need_wsq else None
        _lib.check(_lib.lib().afcm_conv_weight_prep(
            _lib.ptr(wc), Co, Ci, kh, float(pre_scale), int(bool(normalize)), _lib.ptr(w_f32), _lib.ptr(w_tc),
            _lib.dtype_code(dtype), _lib.ptr(wsq), _lib.stream_ptr(w.device)))
        if need_f32:
            ent['w_f32'] = w_f32
        if need_tc:
            ent[('w_tc', dtype)] = w_tc
        if need_wsq:
            ent['wsq'] = wsq
    return ent


def conv2d_native(x, w, padding, icoef=None, ocoef=None, pre_scale=1.0, normalize=False, impl=None, out=None,
                  out_dtype=None, bias=None):
    """y[n,o] = ocoef[n,o] * conv(icoef[n,i] * x[n,i], prepared(w))  -- the shared formulation of the
    encoder conv, Conv2dLayer and modulated_conv2d (include/afcm_b200.h).  The tensor-core path accepts
    float32 or float16 activations and writes `out_dtype` (float32 default); the exact path is float32 only.
    `bias` [Co] is added after the ocoef scale (fused in the tensor-core epilogue)."""
    impl = impl or conv_impl
    _lib.require_cuda(x, w)
    L = _lib.lib()
    N, Ci, H, W = x.shape
    Co, Ci2, kh, kw = w.shape
    assert Ci == Ci2 and kh == kw
    use_tc = impl == 'tc' and kh == 3 and padding in (1, 2)
    out_dtype = out_dtype or torch.float32
    if x.dtype != torch.float32 and not (use_tc and x.dtype == torch.float16):
        raise RuntimeError('afcm conv2d: activations must be float32 (or float16 on the tensor-core path)')
    if out_dtype != torch.float32 and not (use_tc and out_dtype == torch.float16):
        raise RuntimeError('afcm conv2d: the result is float32 (or float16 on the tensor-core path)')
    OH, OW = H + 2 * padding - kh + 1, W + 2 * padding - kw + 1
    y = out if out is not None else torch.empty([N, Co, OH, OW], dtype=out_dtype, device=x.device)
    assert y.dtype == out_dtype and y.is_contiguous()
    st = _lib.stream_ptr(x.device)
    ent = prepare_weight(w, pre_scale, normalize, want_tc=use_tc)
    flops = 2.0 * N * Co * Ci * kh * kw * OH * OW
    # planes stored at the pitch W + 2 with zero pad columns (filtered_lrelu.conv_ready_empty), or dense
    pitch = W + 2 if x.stride() == (Ci * H * (W + 2), H * (W + 2), W + 2, 1) else (W if x.is_contiguous() else 0)
    if use_tc and direct_nchw and padding == 2 and x.dtype == torch.float16 and tc_dtype == torch.float16 and W % 2 == 0 \
            and not _prefers_pack(Ci, Co, H, W):
        # no packed copy of the activations -- the kernel's producer warps build the A tiles from the planes.
        if pitch and (H * pitch) % 8 == 0 and x.data_ptr() % 16 == 0:
            rc = _lib.timed('conv2d_tc', flops, lambda: L.afcm_conv2d_tc_nchw(
                _lib.ptr(x), pitch, _lib.ptr(icoef), _lib.ptr(ent[('w_tc', tc_dtype)]), _lib.ptr(ocoef), _lib.ptr(bias), _lib.ptr(y),
                _lib.dtype_code(out_dtype), N, Ci, H, W, Co, st))
            _lib.check(rc, allow_unsupported=True)
            if rc == 0:
                return y
    if not (use_tc and pitch):
        x = x.contiguous()
        pitch = W
    if use_tc:
        plane = int(L.afcm_conv_tc_plane_elems(H, W, Ci))
        xp = torch.empty([N, plane], dtype=tc_dtype, device=x.device)
        code = _lib.dtype_code(tc_dtype)
        _lib.timed('conv_tc_pack', float(x.element_size() * x.numel() + 2 * xp.numel()), lambda: _lib.check(
            L.afcm_conv_tc_pack_pitched(_lib.ptr(x), _lib.dtype_code(x.dtype), pitch, _lib.ptr(icoef), _lib.ptr(xp), code, N, Ci, H, W, st)))
        _lib.timed('conv2d_tc', flops, lambda: _lib.check(
            L.afcm_conv2d_tc(_lib.ptr(xp), _lib.ptr(ent[('w_tc', tc_dtype)]), _lib.ptr(ocoef), _lib.ptr(bias), _lib.ptr(y),
                             _lib.dtype_code(out_dtype), code, N, Ci, H, W, Co, padding, st)))
        bias = None
    else:
        _lib.check(L.afcm_conv2d_f32(_lib.ptr(x), _lib.ptr(ent['w_f32']), _lib.ptr(icoef), _lib.ptr(ocoef), _lib.ptr(y),
                                     N, Ci, H, W, Co, kh, padding, st))
    if bias is not None:
        y += bias.reshape(1, -1, 1, 1).to(y.dtype)
    return y


class _ConvFn(torch.autograd.Function):
    """Differentiable  y[n,o] = ocoef[n,o] * conv(icoef[n,i] * x[n,i], wp)  (include/afcm_b200.h "convolution
    gradients").  `wp` is the PREPARED weight (scaled / RMS-normalised by the caller with differentiable tensor
    operations on the small weight tensor), icoef / ocoef are [N,Ci] / [N,Co] tensors or None.  This is the backward
    the reference obtains from autograd of F.conv2d through conv2d_gradfix.conv2d (OPS/conv2d_gradfix.py:37-40)
    inside modulated_conv2d (NET:46-63): the data gradient is the forward kernel on flipped, transposed weights,
    the weight gradient one GEMM over the whole batch (no per-sample [N,O,I,k,k] weight gradient exists)."""

    @staticmethod
    def forward(ctx, x, wp, icoef, ocoef, padding, impl):
        _lib.require_cuda(x, wp)
        L = _lib.lib()
        x = x.contiguous().float()
        wp = wp.contiguous().float()
        icoef = None if icoef is None else icoef.contiguous().float()
        ocoef = None if ocoef is None else ocoef.contiguous().float()
        N, Ci, H, W = x.shape
        Co, Ci2, k, k2 = wp.shape
        assert Ci == Ci2 and k == k2 and k in (1, 3)
        OH, OW = H + 2 * padding - k + 1, W + 2 * padding - k + 1
        use_tc = impl == 'tc' and k == 3 and 0 <= padding <= 2
        st = _lib.stream_ptr(x.device)
        y = torch.empty([N, Co, OH, OW], dtype=torch.float32, device=x.device)
        xp = None
        if use_tc:
            xp = _pack(x, icoef, tc_dtype)
            w_tc = _weight_tc(wp, tc_dtype)
            _lib.timed('conv2d_tc', 2.0 * N * Co * Ci * 9 * OH * OW, lambda: _lib.check(
                L.afcm_conv2d_tc(_lib.ptr(xp), _lib.ptr(w_tc), _lib.ptr(ocoef), None, _lib.ptr(y), _lib.F32,
                                 _lib.dtype_code(tc_dtype), N, Ci, H, W, Co, padding, st)))
        else:
            _lib.check(L.afcm_conv2d_f32(_lib.ptr(x), _lib.ptr(wp), _lib.ptr(icoef), _lib.ptr(ocoef), _lib.ptr(y),
                                         N, Ci, H, W, Co, k, padding, st))
        ctx.padding, ctx.use_tc, ctx.tc_dtype = padding, use_tc, tc_dtype
        need_y = ocoef is not None and ctx.needs_input_grad[3]
        need_x = (not use_tc) or (icoef is not None and ctx.needs_input_grad[2]) or (icoef is None and ocoef is None and ctx.needs_input_grad[1])
        ctx.save_for_backward(x if need_x else None, wp, icoef, ocoef, y if need_y else None, xp)
        ctx.xshape = (N, Ci, H, W)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, wp, icoef, ocoef, y, xp = ctx.saved_tensors
        if torch.is_grad_enabled() and icoef is None and ocoef is None:
            # backward under create_graph=True (the R1 penalty of the discriminator step, models/comodgan_model.py:128-161):
            # the gradients are themselves built from differentiable Functions, like the reference's conv2d_gradfix
            # (OPS/conv2d_gradfix.py:83-198 of the upstream op): dx = conv(dy, flip(w)^T), dw = Wgrad(dy, x)
            k, pad = wp.shape[2], ctx.padding
            impl = 'tc' if ctx.use_tc else 'f32'
            dx = dw = None
            if ctx.needs_input_grad[0]:
                dx = _ConvFn.apply(dy, wp.flip(2, 3).transpose(0, 1), None, None, k - 1 - pad, impl)
            if ctx.needs_input_grad[1]:
                xs = x if x is not None else _unpack_like(ctx)
                dw = _WgradFn.apply(dy, xs, pad, k, impl)
            return dx, dw, None, None, None, None
        L = _lib.lib()
        dy = dy.contiguous().float()
        N, Ci, H, W = ctx.xshape
        Co, _, k, _ = wp.shape
        pad = ctx.padding
        OH, OW = dy.shape[2], dy.shape[3]
        st = _lib.stream_ptr(dy.device)
        need_x, need_w, need_ic, need_oc = ctx.needs_input_grad[:4]
        need_ic = need_ic and icoef is not None
        need_oc = need_oc and ocoef is not None
        dx = dw = d_icoef = d_ocoef = None
        if need_oc:                                   # d_ocoef[n,o] = <dy, conv result before the scale> = <dy, y> / ocoef
            d_ocoef = torch.empty_like(ocoef)
            _lib.check(L.afcm_plane_dot_scale(_lib.ptr(dy), _lib.ptr(y), _lib.ptr(ocoef), None, _lib.ptr(d_ocoef),
                                              N * Co, OH * OW, st))
        dyp = None
        code = _lib.dtype_code(ctx.tc_dtype)
        if ctx.use_tc and (need_x or need_ic or need_w):
            dyp = _pack(dy, ocoef, ctx.tc_dtype)      # ocoef folded while packing: serves both gradients
        if need_x or need_ic:
            wT = wp.flip(2, 3).transpose(0, 1).contiguous()                # [Ci, Co, k, k]
            dx = torch.empty([N, Ci, H, W], dtype=torch.float32, device=dy.device)
            if ctx.use_tc:
                w_tc = _weight_tc(wT, ctx.tc_dtype)
                _lib.timed('conv2d_tc_dgrad', 2.0 * N * Co * Ci * 9 * H * W, lambda: _lib.check(
                    L.afcm_conv2d_tc(_lib.ptr(dyp), _lib.ptr(w_tc), None, None, _lib.ptr(dx), _lib.F32, code,
                                     N, Co, OH, OW, Ci, k - 1 - pad, st)))
            else:
                _lib.check(L.afcm_conv2d_f32(_lib.ptr(dy), _lib.ptr(wT), _lib.ptr(ocoef), None, _lib.ptr(dx),
                                             N, Co, OH, OW, Ci, k, k - 1 - pad, st))
            if icoef is not None:                     # d_icoef = <dxm, x>, then dx = icoef * dxm (one launch)
                if need_ic:
                    d_icoef = torch.empty_like(icoef)
                _lib.check(L.afcm_plane_dot_scale(_lib.ptr(dx), _lib.ptr(x), None, _lib.ptr(icoef), _lib.ptr(d_icoef),
                                                  N * Ci, H * W, st))
            if not need_x:
                dx = None
        if need_w:
            dw = torch.empty_like(wp)
            if ctx.use_tc:
                if wgrad_impl == 'tcgen05':
                    nbytes = int(L.afcm_conv2d_wgrad_tc_workspace(N, Ci, H, W, Co, pad))
                    ws = torch.empty([nbytes // 4], dtype=torch.float32, device=dy.device)
                    _lib.timed('conv2d_wgrad_tc', 2.0 * N * Co * Ci * 9 * OH * OW, lambda: _lib.check(
                        L.afcm_conv2d_wgrad_tc5(_lib.ptr(dyp), _lib.ptr(xp), _lib.ptr(dw), _lib.ptr(ws), nbytes, code,
                                                N, Ci, H, W, Co, pad, st)))
                else:
                    _lib.timed('conv2d_wgrad_tc', 2.0 * N * Co * Ci * 9 * OH * OW, lambda: _lib.check(
                        L.afcm_conv2d_wgrad_tc(_lib.ptr(dyp), _lib.ptr(xp), _lib.ptr(dw), code, N, Ci, H, W, Co, pad, st)))
            else:
                _lib.check(L.afcm_conv2d_wgrad_f32(_lib.ptr(dy), _lib.ptr(x), _lib.ptr(icoef), _lib.ptr(ocoef), _lib.ptr(dw),
                                                   N, Ci, H, W, Co, k, pad, st))
        return dx, dw, d_icoef, d_ocoef, None, None


def _unpack_like(ctx):
    raise RuntimeError('afcm conv2d: second-order gradients need the saved fp32 input (run the convolution with impl="f32" or keep x)')


class _WgradFn(torch.autograd.Function):
    """dw[o,i,ky,kx] = sum_{n,p} dy[n,o,p] x[n,i,p + (ky - pad, kx - pad)] as a differentiable Function (needed only when the
    backward pass itself is differentiated: R1).  Its own gradients are convolutions again:
        d(dy) = conv(x, g, pad),      d(x) = conv(dy, flip(g)^T, k - 1 - pad)      for g = grad of the loss w.r.t. dw."""

    @staticmethod
    def forward(ctx, dy, x, pad, k, impl):
        L = _lib.lib()
        dy = dy.contiguous().float()
        x = x.contiguous().float()
        N, Ci, H, W = x.shape
        Co = dy.shape[1]
        dw = torch.empty([Co, Ci, k, k], dtype=torch.float32, device=x.device)
        _lib.check(L.afcm_conv2d_wgrad_f32(_lib.ptr(dy), _lib.ptr(x), None, None, _lib.ptr(dw), N, Ci, H, W, Co, k, pad,
                                           _lib.stream_ptr(x.device)))
        ctx.save_for_backward(dy, x)
        ctx.pad, ctx.k, ctx.impl = pad, k, impl
        return dw

    @staticmethod
    def backward(ctx, g):
        dy, x = ctx.saved_tensors
        pad, k, impl = ctx.pad, ctx.k, ctx.impl
        d_dy = d_x = None
        if ctx.needs_input_grad[0]:
            d_dy = _ConvFn.apply(x, g, None, None, pad, impl)
        if ctx.needs_input_grad[1]:
            d_x = _ConvFn.apply(dy, g.flip(2, 3).transpose(0, 1), None, None, k - 1 - pad, impl)
        return d_dy, d_x, None, None, None


def _pack(x, coef, dtype):
    """NCHW float32 -> 16-bit channel-innermost flat planes with the per-(n,channel) coefficient folded in."""
    L = _lib.lib()
    N, C, H, W = x.shape
    xp = torch.empty([N, int(L.afcm_conv_tc_plane_elems(H, W, C))], dtype=dtype, device=x.device)
    _lib.timed('conv_tc_pack', float(4 * x.numel() + 2 * xp.numel()), lambda: _lib.check(
        L.afcm_conv_tc_pack(_lib.ptr(x), _lib.F32, _lib.ptr(coef), _lib.ptr(xp), _lib.dtype_code(dtype), N, C, H, W,
                            _lib.stream_ptr(x.device))))
    return xp


def _weight_tc(wp, dtype):
    """Prepared float32 weight [Co,Ci,3,3] -> the tap-major 16-bit tile source of the tcgen05 kernel."""
    Co, Ci, k, _ = wp.shape
    w_tc = torch.empty([k * k, (Co + 15) // 16 * 16, (Ci + 63) // 64 * 64], dtype=dtype, device=wp.device)
    _lib.check(_lib.lib().afcm_conv_weight_prep(_lib.ptr(wp), Co, Ci, k, 1.0, 0, None, _lib.ptr(w_tc), _lib.dtype_code(dtype),
                                                None, _lib.stream_ptr(wp.device)))
    return w_tc


def needs_grad(*tensors):
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


def conv2d_train(x, w, padding, icoef=None, ocoef=None, pre_scale=1.0, normalize=False, impl=None):
    """The differentiable counterpart of conv2d_native: weight preparation as tensor operations on the (small) weight
    so that autograd carries it, then _ConvFn."""
    wp = w.float()
    if pre_scale != 1.0:
        wp = wp * float(pre_scale)
    if normalize:
        wp = wp * wp.square().mean([1, 2, 3], keepdim=True).rsqrt()          # NET:42
    return _ConvFn.apply(x, wp, icoef, ocoef, int(padding), impl or conv_impl)


def _pair(v):
    if isinstance(v, (tuple, list)):
        assert len(v) == 2
        return int(v[0]), int(v[1])
    return int(v), int(v)


def _conv_unit(x, w, pad):
    """Stride-1, groups-1 convolution on the native kernels with a symmetric integer padding."""
    if needs_grad(x, w):
        return conv2d_train(x, w, int(pad))
    return conv2d_native(x, w, int(pad))


def conv2d(input, weight, bias=None, stride=1, padding=0, dilation=1, groups=1):
    """Drop-in for conv2d_gradfix.conv2d (OPS/conv2d_gradfix.py:37-40, CM/torch_utils/ops/conv2d_gradfix.py): square 1x1 / 3x3
    kernels, dilation 1, any stride, (py, px) padding and group count.  The generator path (stride 1, groups 1, symmetric
    padding) is ONE native kernel; the other forms, used by the discriminator and the CoModGAN baseline generator (CM/layers.py,
    CM/generator.py:614-836 through conv2d_resample), are compositions of native kernels: unequal padding = zero padding with
    upfirdn2d, stride s = stride-1 convolution followed by keeping every s-th sample (upfirdn2d down), groups = one convolution
    per group.  Differentiable (first order) with respect to input, weight and bias."""
    from . import upfirdn2d
    assert isinstance(input, torch.Tensor)
    if _pair(dilation) != (1, 1):
        raise NotImplementedError('afcm conv2d supports dilation 1 only')
    sy, sx = _pair(stride)
    py, px = _pair(padding)
    groups = int(groups)
    if groups > 1:
        N, C, H, W = input.shape
        O = weight.shape[0]
        assert C % groups == 0 and O % groups == 0 and weight.shape[1] == C // groups
        ci, co = C // groups, O // groups
        ys = [conv2d(input[:, g * ci:(g + 1) * ci].contiguous(), weight[g * co:(g + 1) * co].contiguous(), None, (sy, sx), (py, px))
              for g in range(groups)]
        y = torch.cat(ys, dim=1)
    else:
        x = input
        if py != px:
            x = upfirdn2d.upfirdn2d(x, None, padding=[px, px, py, py])          # explicit zero padding (native kernel)
            pad = 0
        else:
            pad = py
        kh = weight.shape[2]
        if pad > kh - 1:                                                        # the native kernels take up to "full" padding
            x = upfirdn2d.upfirdn2d(x, None, padding=[pad - (kh - 1)] * 4)
            pad = kh - 1
        y = _conv_unit(x, weight, pad)
        if (sy, sx) != (1, 1):
            y = upfirdn2d.upfirdn2d(y, None, down=[sx, sy])                      # keep every s-th sample
    if bias is not None:
        y = y + bias.reshape(1, -1, 1, 1)
    return y


def conv_transpose2d(input, weight, bias=None, stride=1, padding=0, output_padding=0, groups=1, dilation=1):
    """Drop-in for conv2d_gradfix.conv_transpose2d (OPS/conv2d_gradfix.py:42, used by conv2d_resample for up-sampling layers,
    CM/torch_utils/ops/conv2d_resample.py:127-141): weight [Cin, Cout / groups, kh, kw].  Evaluated as the stride-1 convolution
    of the zero-inserted input with the flipped, transposed weight -- zero insertion and the k-1-p border by the native upfirdn2d
    kernel, the convolution by the native convolution kernels."""
    from . import upfirdn2d
    if _pair(dilation) != (1, 1) or _pair(output_padding) != (0, 0):
        raise NotImplementedError('afcm conv_transpose2d supports dilation 1 and output_padding 0 only')
    sy, sx = _pair(stride)
    py, px = _pair(padding)
    groups = int(groups)
    Cin, Cog, kh, kw = weight.shape
    assert input.shape[1] == Cin and Cin % groups == 0
    if groups > 1:
        cig = Cin // groups
        ys = [conv_transpose2d(input[:, g * cig:(g + 1) * cig].contiguous(), weight[g * cig:(g + 1) * cig].contiguous(), None, (sy, sx), (py, px))
              for g in range(groups)]
        y = torch.cat(ys, dim=1)
    else:
        # zero insertion leaves s - 1 zeros behind the last sample: the transposed convolution's grid ends at the sample itself
        bx0, bx1 = kw - 1 - px, kw - 1 - px - (sx - 1)
        by0, by1 = kh - 1 - py, kh - 1 - py - (sy - 1)
        x = upfirdn2d.upfirdn2d(input, None, up=[sx, sy], padding=[bx0, bx1, by0, by1])
        w = weight.flip([2, 3]).transpose(0, 1).contiguous()                    # [Cout, Cin, kh, kw], correlation form
        y = _conv_unit(x, w, 0)
    if bias is not None:
        y = y + bias.reshape(1, -1, 1, 1)
    return y
