"""filtered_lrelu -- same public signature as the reference op
(models/networks/stylegan3/torch_utils/ops/filtered_lrelu.py:56), executed by the hand-written sm_100a
kernel behind afcm_filtered_lrelu (include/afcm_b200.h).  `impl` is accepted for compatibility and
ignored: there is one implementation and it needs a CUDA tensor."""
import numpy as np
import torch

from ... import _lib
from . import upfirdn2d as _upfirdn2d


# Implementation of the autograd (training) path: 'exact' = the fp32 kernel (afcm_filtered_lrelu, the parity path),
# 'tcs' = the shared-memory-tiled tensor-core kernel with the same sign tensor (afcm_filtered_lrelu_tcs: fp16 operands in the
# forward, bf16 in the backward, fp32 accumulation and activation), 'tc' = the register-chained tensor-core kernel with the sign
# tensor (afcm_filtered_lrelu_tc_signs: fp16 operands, the backward scaled into the fp16 range by max|dy|; falls back to 'tcs' /
# 'exact' for calls it has no specialisation for).  All three write and read the SAME sign tensor.  Inference is not affected.
train_impl = 'exact'


def set_train_impl(impl):
    global train_impl
    assert impl in ('exact', 'tc', 'tcs')
    train_impl = impl


_amax_buf = {}
tc_sign_calls = 0            # calls served by afcm_filtered_lrelu_tc_signs (tests check that the kernel really ran)


def _absmax(x):
    """max|x| as a device scalar (no host synchronisation): the operand scale of the fp16 backward."""
    buf = _amax_buf.get(x.device)
    if buf is None:
        buf = _amax_buf[x.device] = torch.zeros([1], dtype=torch.float32, device=x.device)
    _lib.check(_lib.lib().afcm_absmax(_lib.ptr(x), x.numel(), _lib.ptr(buf), _lib.stream_ptr(x.device)))
    return buf


def _get_filter_size(f):
    if f is None:
        return 1, 1
    assert isinstance(f, torch.Tensor) and 1 <= f.ndim <= 2
    return f.shape[-1], f.shape[0]  # width, height


def _parse_padding(padding):
    # same accepted forms as the reference (filtered_lrelu.py:42-52)
    if isinstance(padding, (int, np.integer)):
        padding = [padding, padding]
    assert isinstance(padding, (list, tuple))
    assert all(isinstance(x, (int, np.integer)) for x in padding)
    padding = [int(x) for x in padding]
    if len(padding) == 2:
        px, py = padding
        padding = [px, px, py, py]
    px0, px1, py0, py1 = padding
    return px0, px1, py0, py1


def filtered_lrelu(x, fu=None, fd=None, b=None, up=1, down=1, padding=0, gain=np.sqrt(2), slope=0.2, clamp=None,
                   flip_filter=False, impl='cuda'):
    r"""Filtered leaky ReLU: bias, zero-insert upsample by `up`, pad, FIR `fu`, gain, leaky ReLU, clamp,
    FIR `fd`, keep every `down`-th sample.  Arguments as in the reference (filtered_lrelu.py:85-111).
    Supports first-order gradients w.r.t. `x` and `b` (the backward pass re-runs the op with the filters
    swapped, reading the 2-bit sign tensor written by the forward pass, like the reference)."""
    assert isinstance(x, torch.Tensor)
    assert impl in ['ref', 'cuda']
    _lib.require_cuda(x)
    return _filtered_lrelu_cuda(up=up, down=down, padding=padding, gain=gain, slope=slope, clamp=clamp,
                                flip_filter=flip_filter).apply(x, fu, fd, b, None, 0, 0)


def _taps_1d(f):
    """(host taps or None, taps) for a separable filter; None for a 2-D filter (no fused kernel)."""
    if f is None:
        return None, 1
    if f.ndim == 1:
        return _lib.host_array(f), int(f.shape[0])
    if f.ndim == 2 and f.shape[0] == 1 and f.shape[1] == 1:
        return _lib.host_array(f).reshape(1), 1
    return False, 0


def _run_fused(x, fu, fd, b, si, up, down, px0, px1, py0, py1, sx, sy, gain, slope, clamp, flip_filter, write_signs,
               skip=None, out_scale=1.0):
    """Calls afcm_filtered_lrelu.  Returns (y, so, return_code) like the reference plugin
    (filtered_lrelu.cpp:16-18,208): return_code -1 means 'no fused kernel for this geometry'."""
    L = _lib.lib()
    fu_h, fu_n = _taps_1d(fu)
    fd_h, fd_n = _taps_1d(fd)
    if fu_h is False or fd_h is False or x.dtype not in (torch.float32, torch.float16):
        return None, None, -1
    N, C, xh, xw = x.shape
    yh, yw = _lib._c.c_int(), _lib._c.c_int()
    _lib.check(L.afcm_filtered_lrelu_out_size(xh, xw, up, down, fu_n, fd_n, px0, px1, py0, py1, yh, yw))
    yh, yw = yh.value, yw.value
    y = torch.empty([N, C, yh, yw], dtype=x.dtype, device=x.device)
    so = None
    mode, s, sh, swb = _lib.SIGN_NONE, None, 0, 0
    if write_signs:
        a, c = _lib._c.c_int(), _lib._c.c_int()
        L.afcm_filtered_lrelu_sign_size(yh, yw, down, fd_n, a, c)
        sh, swb = a.value, c.value
        so = torch.empty([N, C, sh, swb], dtype=torch.uint8, device=x.device)
        mode, s = _lib.SIGN_WRITE, so
    elif si is not None and si.numel():
        assert si.is_contiguous() and si.dtype == torch.uint8 and si.ndim == 4
        mode, s, sh, swb = _lib.SIGN_READ, si, si.shape[2], si.shape[3]
    if b is not None:
        b = b.contiguous()
    if skip is not None:
        assert skip.shape == y.shape and skip.dtype == y.dtype
        skip = skip.contiguous()
    # algorithmic bytes: x read once + y written once (+ the skip read when fused)
    nbytes = x.element_size() * (x.numel() + y.numel() * (2 if skip is not None else 1))
    tc_geo = (mode != _lib.SIGN_NONE and x.dtype == torch.float32 and fu_h is not None and fd_h is not None and
              (up, down) in ((2, 2), (2, 4), (4, 2)) and fu_n == 6 * up and fd_n == 6 * down)
    # measured per geometry at batch 32 (tools/flr_train_bench.py): the register-chained kernel wins the forward everywhere and
    # the up 2 / down 2 backward; the backward kernels with a 24-tap filter (up 4 / down 2 and up 2 / down 4: 6 column blocks per
    # warp, two CTAs per SM) lose to the shared-memory tiled kernel on planes it tiles well
    tcs_better = mode == _lib.SIGN_READ and (up, down) != (2, 2) and min(yh, yw) >= 48
    if (train_impl == 'tc' and tc_geo and not tcs_better and skip is None and out_scale == 1.0 and xw % 2 == 0 and yw % 2 == 0 and
            (mode == _lib.SIGN_READ or (1.0 / 1024 <= clamp <= 1024))):
        # the register-chained kernel with the sign tensor (csrc/flr_tc.cu, SIGN = 1 / 2)
        xc = x if (x.stride(3) == 1 and all(int(v) % 2 == 0 for v in x.stride()[:3]) and x.data_ptr() % 8 == 0) else x.contiguous()
        amax = _absmax(xc if xc.is_contiguous() else xc.contiguous()) if mode == _lib.SIGN_READ else None
        rc = _lib.timed('filtered_lrelu', nbytes, lambda: L.afcm_filtered_lrelu_tc_signs(
            _lib.ptr(xc), _lib.i64x4(xc.stride()), _lib.ptr(y), _lib.i64x4(y.stride()), _lib.ptr(b),
            N, C, xh, xw, yh, yw, _lib.np_ptr(fu_h), fu_n, _lib.np_ptr(fd_h), fd_n, up, down, px0, px1, py0, py1,
            gain, slope, clamp, int(bool(flip_filter)), mode, _lib.ptr(s), sh, swb, int(sx), int(sy), _lib.ptr(amax),
            _lib.stream_ptr(x.device)))
        _lib.check(rc, allow_unsupported=True)
        if rc == 0:
            global tc_sign_calls
            tc_sign_calls += 1
            return y, so, 0
    use_tc = (train_impl in ('tc', 'tcs') and tc_geo and
              min(yh, yw) >= 48)      # its 32x32 / 16x16 tiles waste most of a small plane: the exact kernel is faster there
    if use_tc:
        # forward (sign write): fp16 operands; backward (sign read): bf16 operands keep the gradients' exponent range
        op = _lib.F16 if mode == _lib.SIGN_WRITE else _lib.BF16
        rc = _lib.timed('filtered_lrelu', nbytes, lambda: L.afcm_filtered_lrelu_tcs(
            _lib.ptr(x), _lib.i64x4(x.stride()), _lib.ptr(y), _lib.i64x4(y.stride()), _lib.ptr(b), _lib.ptr(skip), op,
            N, C, xh, xw, yh, yw, _lib.np_ptr(fu_h), fu_n, _lib.np_ptr(fd_h), fd_n, up, down, px0, px1, py0, py1,
            gain, slope, clamp, out_scale, int(bool(flip_filter)), mode, _lib.ptr(s), sh, swb, int(sx), int(sy),
            _lib.stream_ptr(x.device)))
        _lib.check(rc)
        return y, so, 0
    rc = _lib.timed('filtered_lrelu', nbytes, lambda: L.afcm_filtered_lrelu(
        _lib.ptr(x), _lib.i64x4(x.stride()), _lib.ptr(y), _lib.i64x4(y.stride()), _lib.ptr(b), _lib.ptr(skip),
        _lib.dtype_code(x.dtype), N, C, xh, xw, yh, yw,
        _lib.np_ptr(fu_h), fu_n, _lib.np_ptr(fd_h), fd_n, up, down, px0, px1, py0, py1,
        gain, slope, clamp, out_scale, int(bool(flip_filter)), mode, _lib.ptr(s), sh, swb, int(sx), int(sy),
        _lib.stream_ptr(x.device)))
    _lib.check(rc, allow_unsupported=True)
    if rc != 0:
        return None, None, rc
    return y, so, 0


# The tcgen05 / TMEM kernel (csrc/flr_t5.cu) runs when asked for by name (impl='t5') or, with t5_enabled = True, whenever the
# call qualifies (fp16 input with 16-byte aligned strides, no bias).  It is off by default: measured on B200 it is latency-chain
# bound at 8.5 % of HBM over the 28 layers against 25.7 % for the register-chained mma.sync kernel.
t5_enabled = False


def padded_pitch_empty(shape, dtype, device, align_bytes=16):
    """An [N,C,H,W] tensor whose row pitch is padded to `align_bytes` (the TMA input requirement of the tcgen05
    filtered_lrelu: every stride a multiple of 16 bytes).  Returns the [N,C,H,W] view of the padded allocation."""
    N, C, H, W = [int(v) for v in shape]
    q = max(1, align_bytes // torch.empty([], dtype=dtype).element_size())
    Wp = (W + q - 1) // q * q
    return torch.empty([N, C, H, Wp], dtype=dtype, device=device)[..., :W]


def t5_eligible(x, b):
    """Whether afcm_filtered_lrelu_t5 accepts this input (the library re-checks and answers AFCM_ERR_UNSUPPORTED)."""
    if not t5_enabled or b is not None or x.dtype != torch.float16 or x.stride(3) != 1:
        return False
    return x.data_ptr() % 16 == 0 and all((int(st) * 2) % 16 == 0 for st in x.stride()[:3])


def conv_ready_empty(shape, dtype, device):
    """An [N,C,H,W] result tensor stored at the row pitch W + 2: with `zero_pad_cols=2` filtered_lrelu_tc() writes zeros into the
    two extra columns, which makes the allocation the flat plane of row pitch W + 2 that the tcgen05 convolution reads directly
    (conv2d_gradfix.conv2d_native, afcm_conv2d_tc_nchw with x_pitch = W + 2).  Returns the [N,C,H,W] view."""
    N, C, H, W = [int(v) for v in shape]
    return torch.empty([N, C, H, W + 2], dtype=dtype, device=device)[..., :W]


def filtered_lrelu_tc(x, fu, fd, b=None, up=1, down=1, padding=0, gain=np.sqrt(2), slope=0.2, clamp=None,
                      flip_filter=False, skip=None, out_scale=1.0, out_dtype=None, out=None, impl=None, zero_pad_cols=0, conv_ready=False):
    """Tensor-core forward of filtered_lrelu (afcm_filtered_lrelu_tc, csrc/flr_tc.cu): same arguments and
    result as filtered_lrelu() up to fp16 operand rounding (max |err| <= 2e-3 * max|y|).  Inference only (no
    autograd, no sign tensor).  `skip` is added to the result and `out_scale` multiplies it (NET:376-377,
    NET:699-700 fused); x may be float32 or float16, `out_dtype` selects the result type.  Returns None when
    the geometry has no tensor-core kernel (the caller then uses filtered_lrelu())."""
    _lib.require_cuda(x, fu, fd, b, skip)
    L = _lib.lib()
    px0, px1, py0, py1 = _parse_padding(padding)
    fu_h, fu_n = _taps_1d(fu)
    fd_h, fd_n = _taps_1d(fd)
    if fu_h is False or fd_h is False or fu_h is None or fd_h is None:
        return None
    if x.dtype not in (torch.float32, torch.float16) or x.stride(3) != 1:
        return None
    if not ((up, down) in ((2, 2), (4, 2), (2, 4)) and fu_n == 6 * up and fd_n == 6 * down):
        return None
    out_dtype = out_dtype or x.dtype
    N, C, xh, xw = x.shape
    yh, yw = _lib._c.c_int(), _lib._c.c_int()
    _lib.check(L.afcm_filtered_lrelu_out_size(xh, xw, up, down, fu_n, fd_n, px0, px1, py0, py1, yh, yw))
    yh, yw = yh.value, yw.value
    if conv_ready and out is None and impl != 't5' and yw % 2 == 0:
        y = conv_ready_empty([N, C, yh, yw], out_dtype, x.device)         # row pitch yw + 2, the kernel writes the zero columns
        zero_pad_cols = 2
    else:
        y = out if out is not None else torch.empty([N, C, yh, yw], dtype=out_dtype, device=x.device)
    assert y.shape == (N, C, yh, yw) and y.dtype == out_dtype and y.stride(3) == 1
    if b is not None:
        b = b.detach().float().contiguous()
    if skip is not None:
        assert skip.shape == y.shape and skip.dtype == y.dtype
        if skip.stride() != y.stride():
            skip = skip.contiguous() if y.is_contiguous() else conv_ready_empty(y.shape, y.dtype, y.device).copy_(skip)
    clamp = float(clamp) if clamp is not None else float('inf')
    nbytes = x.element_size() * x.numel() + y.element_size() * y.numel() * (2 if skip is not None else 1)
    if impl == 't5' or (impl is None and t5_eligible(x, b)):
        # tcgen05 / TMEM kernel; AFCM_ERR_UNSUPPORTED (geometry, clamp range, strides) falls through to the mma.sync kernel
        rc = _lib.timed('filtered_lrelu', nbytes, lambda: L.afcm_filtered_lrelu_t5(
            _lib.ptr(x), _lib.i64x4(x.stride()), _lib.dtype_code(x.dtype), _lib.ptr(y), _lib.i64x4(y.stride()),
            _lib.dtype_code(y.dtype), _lib.ptr(b), _lib.ptr(skip), N, C, xh, xw, yh, yw,
            _lib.np_ptr(fu_h), fu_n, _lib.np_ptr(fd_h), fd_n, up, down, px0, px1, py0, py1,
            float(gain), float(slope), clamp, float(out_scale), int(bool(flip_filter)), _lib.stream_ptr(x.device)))
        _lib.check(rc, allow_unsupported=True)
        if rc == 0:
            return y
        if impl == 't5':
            return None
    rc = _lib.timed('filtered_lrelu', nbytes, lambda: L.afcm_filtered_lrelu_tc_padded(
        _lib.ptr(x), _lib.i64x4(x.stride()), _lib.dtype_code(x.dtype), _lib.ptr(y), _lib.i64x4(y.stride()),
        _lib.dtype_code(y.dtype), _lib.ptr(b), _lib.ptr(skip), N, C, xh, xw, yh, yw,
        _lib.np_ptr(fu_h), fu_n, _lib.np_ptr(fd_h), fd_n, up, down, px0, px1, py0, py1,
        float(gain), float(slope), clamp, float(out_scale), int(bool(flip_filter)), int(zero_pad_cols), _lib.stream_ptr(x.device)))
    _lib.check(rc, allow_unsupported=True)
    return y if rc == 0 else None


def _act_(y, si, sx, sy, gain, slope, clamp, write_signs):
    """In-place activation step of the generic composition (reference: filtered_lrelu_act_)."""
    L = _lib.lib()
    assert y.is_contiguous()
    N, C, h, w = y.shape
    so = None
    mode, s, sh, swb = _lib.SIGN_NONE, None, 0, 0
    if write_signs:
        swb = ((w + 15) & ~15) >> 2
        sh = h
        so = torch.empty([N, C, sh, swb], dtype=torch.uint8, device=y.device)
        mode, s = _lib.SIGN_WRITE, so
    elif si is not None and si.numel():
        mode, s, sh, swb = _lib.SIGN_READ, si, si.shape[2], si.shape[3]
    _lib.check(L.afcm_filtered_lrelu_act(_lib.ptr(y), _lib.dtype_code(y.dtype), N * C, h, w, gain, slope, clamp,
                                         mode, _lib.ptr(s), sh, swb, int(sx), int(sy), _lib.stream_ptr(y.device)))
    return so


def _channel_sum(dx):
    """db = dx.sum([0, 2, 3]) (reference filtered_lrelu.py:264-265): native per-plane sums, then the tiny sum over N."""
    if dx.dtype != torch.float32 or not dx.is_contiguous() or dx.numel() == 0:
        return dx.sum([0, 2, 3])
    N, C, H, W = dx.shape
    part = torch.empty([N, C], dtype=torch.float32, device=dx.device)
    _lib.check(_lib.lib().afcm_plane_sum(_lib.ptr(dx), _lib.ptr(part), N * C, H * W, _lib.stream_ptr(dx.device)))
    return part.sum(0)


_filtered_lrelu_cuda_cache = dict()


def _filtered_lrelu_cuda(up=1, down=1, padding=0, gain=np.sqrt(2), slope=0.2, clamp=None, flip_filter=False):
    assert isinstance(up, (int, np.integer)) and up >= 1
    assert isinstance(down, (int, np.integer)) and down >= 1
    up, down = int(up), int(down)
    px0, px1, py0, py1 = _parse_padding(padding)
    assert gain == float(gain) and gain > 0
    gain = float(gain)
    assert slope == float(slope) and slope >= 0
    slope = float(slope)
    assert clamp is None or (clamp == float(clamp) and clamp >= 0)
    clamp = float(clamp if clamp is not None else 'inf')

    key = (up, down, px0, px1, py0, py1, gain, slope, clamp, flip_filter)
    if key in _filtered_lrelu_cuda_cache:
        return _filtered_lrelu_cuda_cache[key]

    class FilteredLReluCuda(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, fu, fd, b, si, sx, sy):
            assert isinstance(x, torch.Tensor) and x.ndim == 4
            _lib.require_cuda(x, fu, fd, b)
            if b is not None:
                assert b.dtype == x.dtype and b.ndim == 1 and b.shape[0] == x.shape[1]
            write_signs = (si is None or si.numel() == 0) and (x.requires_grad or (b is not None and b.requires_grad))
            y, so, rc = _run_fused(x, fu, fd, b, si, up, down, px0, px1, py0, py1, sx, sy, gain, slope, clamp,
                                   flip_filter, write_signs)
            if rc < 0:
                # Generic composition, same as the reference's fallback (filtered_lrelu.py:223-229):
                # bias, upfirdn2d(up), in-place activation with sign handling, upfirdn2d(down).
                y = x if b is None else x + b.reshape(1, -1, 1, 1)
                y = _upfirdn2d.upfirdn2d(y, fu, up=up, padding=[px0, px1, py0, py1], gain=up ** 2,
                                         flip_filter=flip_filter).contiguous()
                if y is x:
                    y = y.clone()
                so = _act_(y, si, sx, sy, gain, slope, clamp, write_signs)
                y = _upfirdn2d.upfirdn2d(y, fd, down=down, flip_filter=flip_filter)
            ctx.save_for_backward(fu, fd, (si if (si is not None and si.numel()) else so))
            ctx.x_shape = x.shape
            ctx.y_shape = y.shape
            ctx.s_ofs = sx, sy
            ctx.has_b = b is not None
            return y

        @staticmethod
        def backward(ctx, dy):
            fu, fd, si = ctx.saved_tensors
            _, _, xh, xw = ctx.x_shape
            _, _, yh, yw = ctx.y_shape
            sx, sy = ctx.s_ofs
            dx = None
            db = None
            fuw, fuh = _get_filter_size(fu)
            fdw, fdh = _get_filter_size(fd)
            if ctx.needs_input_grad[0] or ctx.needs_input_grad[3]:
                # reference: filtered_lrelu.py:252-263
                pp = [(fuw - 1) + (fdw - 1) - px0, xw * up - yw * down + px0 - (up - 1),
                      (fuh - 1) + (fdh - 1) - py0, xh * up - yh * down + py0 - (up - 1)]
                gg = gain * (up ** 2) / (down ** 2)
                ff = (not flip_filter)
                sx = sx - (fuw - 1) + px0
                sy = sy - (fuh - 1) + py0
                dx = _filtered_lrelu_cuda(up=down, down=up, padding=pp, gain=gg, slope=slope, clamp=None,
                                          flip_filter=ff).apply(dy, fd, fu, None, si, sx, sy)
            if ctx.needs_input_grad[3] and ctx.has_b:
                db = _channel_sum(dx)
            return dx, None, None, db, None, None, None

    _filtered_lrelu_cuda_cache[key] = FilteredLReluCuda
    return FilteredLReluCuda
