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
import torch

from ... import _lib

conv_impl = 'f32'
tc_dtype = torch.float16
act_dtype = torch.float32
enabled = False                      # kept for API compatibility with the reference module
weight_gradients_disabled = False

_prep_cache = {}


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
    if ent is not None and (ent['_ref']() is not w or ent['_version'] != w._version):
        ent = None                      # address / id reuse or in-place update (optimizer step): re-prepare
    if ent is None:
        if len(_prep_cache) > 512:
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
    x = x.contiguous()
    OH, OW = H + 2 * padding - kh + 1, W + 2 * padding - kw + 1
    y = out if out is not None else torch.empty([N, Co, OH, OW], dtype=out_dtype, device=x.device)
    assert y.dtype == out_dtype and y.is_contiguous()
    st = _lib.stream_ptr(x.device)
    ent = prepare_weight(w, pre_scale, normalize, want_tc=use_tc)
    if use_tc:
        plane = int(L.afcm_conv_tc_plane_elems(H, W, Ci))
        xp = torch.empty([N, plane], dtype=tc_dtype, device=x.device)
        code = _lib.dtype_code(tc_dtype)
        flops = 2.0 * N * Co * Ci * kh * kw * OH * OW
        _lib.timed('conv_tc_pack', float(x.element_size() * x.numel() + 2 * xp.numel()), lambda: _lib.check(
            L.afcm_conv_tc_pack(_lib.ptr(x), _lib.dtype_code(x.dtype), _lib.ptr(icoef), _lib.ptr(xp), code, N, Ci, H, W, st)))
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


def conv2d(input, weight, bias=None, stride=1, padding=0, dilation=1, groups=1):
    """Drop-in for conv2d_gradfix.conv2d (reference :37-40) for the configurations the generator uses:
    stride 1, dilation 1, groups 1, square 1x1 / 3x3 kernels, symmetric integer padding.  Forward only --
    gradients of the convolution are the training-step row of the scope table."""
    assert isinstance(input, torch.Tensor)
    if isinstance(padding, (tuple, list)):
        assert padding[0] == padding[1]
        padding = padding[0]
    if stride not in (1, (1, 1)) or dilation not in (1, (1, 1)) or groups != 1:
        raise NotImplementedError('afcm conv2d supports stride=1, dilation=1, groups=1 only')
    if torch.is_grad_enabled() and (input.requires_grad or weight.requires_grad):
        raise NotImplementedError('afcm conv2d: backward is not implemented yet (forward-only hot path)')
    y = conv2d_native(input, weight, int(padding))
    if bias is not None:
        y = y + bias.reshape(1, -1, 1, 1)
    return y


def conv_transpose2d(*args, **kwargs):
    raise NotImplementedError('conv_transpose2d is not on the AFCM stylegan3 generator path')
