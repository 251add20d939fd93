"""upfirdn2d and its convenience wrappers -- same public API as the reference module
(models/networks/stylegan3/torch_utils/ops/upfirdn2d.py:70,118,277,313,352), executed by
afcm_upfirdn2d (include/afcm_b200.h).  `impl` is accepted and ignored."""
import numpy as np
import torch

from ... import _lib


def _parse_scaling(scaling):
    if isinstance(scaling, (int, np.integer)):
        scaling = [scaling, scaling]
    assert isinstance(scaling, (list, tuple))
    assert all(isinstance(x, (int, np.integer)) for x in scaling)
    sx, sy = [int(v) for v in scaling]
    assert sx >= 1 and sy >= 1
    return sx, sy


def _parse_padding(padding):
    if isinstance(padding, (int, np.integer)):
        padding = [padding, padding]
    assert isinstance(padding, (list, tuple))
    assert all(isinstance(x, (int, np.integer)) for x in padding)
    padding = [int(x) for x in padding]
    if len(padding) == 2:
        padx, pady = padding
        padding = [padx, padx, pady, pady]
    padx0, padx1, pady0, pady1 = padding
    return padx0, padx1, pady0, pady1


def _get_filter_size(f):
    if f is None:
        return 1, 1
    assert isinstance(f, torch.Tensor) and f.ndim in [1, 2]
    fw, fh = int(f.shape[-1]), int(f.shape[0])
    assert fw >= 1 and fh >= 1
    return fw, fh


def setup_filter(f, device=torch.device('cpu'), normalize=True, flip_filter=False, gain=1, separable=None):
    """Builds the float32 FIR filter tensor used by upfirdn2d (reference upfirdn2d.py:70-114):
    list / array / tensor -> [taps] (separable, >= 8 taps by default) or [fh, fw]."""
    if f is None:
        f = 1
    f = torch.as_tensor(f, dtype=torch.float32)
    assert f.ndim in [0, 1, 2]
    assert f.numel() > 0
    if f.ndim == 0:
        f = f[np.newaxis]
    if separable is None:
        separable = (f.ndim == 1 and f.numel() >= 8)
    if f.ndim == 1 and not separable:
        f = f.ger(f)
    assert f.ndim == (1 if separable else 2)
    if normalize:
        f = f / f.sum()
    if flip_filter:
        f = f.flip(list(range(f.ndim)))
    f = f * (gain ** (f.ndim / 2))
    return f.to(device=device)


def _launch(x, f_host, fh, fw, upx, upy, downx, downy, px0, px1, py0, py1, flip, gain):
    N, C, xh, xw = x.shape
    yw = (xw * upx + px0 + px1 - fw + downx) // downx
    yh = (xh * upy + py0 + py1 - fh + downy) // downy
    assert yw >= 1 and yh >= 1
    x = x.contiguous()
    y = torch.empty([N, C, yh, yw], dtype=x.dtype, device=x.device)
    _lib.check(_lib.lib().afcm_upfirdn2d(_lib.ptr(x), _lib.ptr(y), _lib.dtype_code(x.dtype), N * C, xh, xw, yh, yw,
                                         _lib.np_ptr(f_host), fh, fw, upx, upy, downx, downy, px0, px1, py0, py1,
                                         int(bool(flip)), float(gain), _lib.stream_ptr(x.device)))
    return y


def upfirdn2d(x, f, up=1, down=1, padding=0, flip_filter=False, gain=1, impl='cuda'):
    r"""Pad, zero-insert upsample, FIR filter and downsample a batch of 2-D images
    (semantics of the reference upfirdn2d.py:118-162)."""
    assert isinstance(x, torch.Tensor) and x.ndim == 4
    assert impl in ['ref', 'cuda']
    _lib.require_cuda(x, f)
    return _upfirdn2d_cuda(up=up, down=down, padding=padding, flip_filter=flip_filter, gain=gain).apply(x, f)


_upfirdn2d_cuda_cache = dict()


def _upfirdn2d_cuda(up=1, down=1, padding=0, flip_filter=False, gain=1):
    upx, upy = _parse_scaling(up)
    downx, downy = _parse_scaling(down)
    padx0, padx1, pady0, pady1 = _parse_padding(padding)
    key = (upx, upy, downx, downy, padx0, padx1, pady0, pady1, flip_filter, gain)
    if key in _upfirdn2d_cuda_cache:
        return _upfirdn2d_cuda_cache[key]

    class Upfirdn2dCuda(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, f):
            if x.dtype not in (torch.float32, torch.float16):
                raise RuntimeError('upfirdn2d: x must be float16 or float32')
            if f is None:
                y = _launch(x, None, 1, 1, upx, upy, downx, downy, padx0, padx1, pady0, pady1, flip_filter, gain)
                f = torch.ones([1, 1], dtype=torch.float32, device=x.device)
            elif f.ndim == 2:
                fh_, fw_ = f.shape
                y = _launch(x, _lib.host_array(f), fh_, fw_, upx, upy, downx, downy, padx0, padx1, pady0, pady1,
                            flip_filter, gain)
            else:
                # separable: horizontal pass then vertical pass, whole gain in the second (reference :241-245)
                fh_host = _lib.host_array(f)
                n = f.shape[0]
                y = _launch(x, fh_host, 1, n, upx, 1, downx, 1, padx0, padx1, 0, 0, flip_filter, 1.0)
                y = _launch(y, fh_host, n, 1, 1, upy, 1, downy, 0, 0, pady0, pady1, flip_filter, gain)
            ctx.save_for_backward(f)
            ctx.x_shape = x.shape
            return y

        @staticmethod
        def backward(ctx, dy):
            f, = ctx.saved_tensors
            _, _, ih, iw = ctx.x_shape
            _, _, oh, ow = dy.shape
            fw, fh = _get_filter_size(f)
            p = [fw - padx0 - 1, iw * upx - ow * downx + padx0 - upx + 1,
                 fh - pady0 - 1, ih * upy - oh * downy + pady0 - upy + 1]
            dx = None
            if ctx.needs_input_grad[0]:
                dx = _upfirdn2d_cuda(up=[downx, downy], down=[upx, upy], padding=p, flip_filter=(not flip_filter),
                                     gain=gain).apply(dy, f)
            return dx, None

    _upfirdn2d_cuda_cache[key] = Upfirdn2dCuda
    return Upfirdn2dCuda


def filter2d(x, f, padding=0, flip_filter=False, gain=1, impl='cuda'):
    """Same-size FIR filtering (reference upfirdn2d.py:277-309)."""
    padx0, padx1, pady0, pady1 = _parse_padding(padding)
    fw, fh = _get_filter_size(f)
    p = [padx0 + fw // 2, padx1 + (fw - 1) // 2, pady0 + fh // 2, pady1 + (fh - 1) // 2]
    return upfirdn2d(x, f, padding=p, flip_filter=flip_filter, gain=gain, impl=impl)


def upsample2d(x, f, up=2, padding=0, flip_filter=False, gain=1, impl='cuda'):
    """FIR upsampling by an integer factor (reference upfirdn2d.py:313-348)."""
    upx, upy = _parse_scaling(up)
    padx0, padx1, pady0, pady1 = _parse_padding(padding)
    fw, fh = _get_filter_size(f)
    p = [padx0 + (fw + upx - 1) // 2, padx1 + (fw - upx) // 2, pady0 + (fh + upy - 1) // 2, pady1 + (fh - upy) // 2]
    return upfirdn2d(x, f, up=up, padding=p, flip_filter=flip_filter, gain=gain * upx * upy, impl=impl)


def downsample2d(x, f, down=2, padding=0, flip_filter=False, gain=1, impl='cuda'):
    """FIR downsampling by an integer factor (reference upfirdn2d.py:352-387)."""
    downx, downy = _parse_scaling(down)
    padx0, padx1, pady0, pady1 = _parse_padding(padding)
    fw, fh = _get_filter_size(f)
    p = [padx0 + (fw - downx + 1) // 2, padx1 + (fw - downx) // 2, pady0 + (fh - downy + 1) // 2,
         pady1 + (fh - downy) // 2]
    return upfirdn2d(x, f, down=down, padding=p, flip_filter=flip_filter, gain=gain, impl=impl)
