"""bias_act -- same public API as the reference op (models/networks/stylegan3/torch_utils/ops/
bias_act.py:52), executed by afcm_bias_act.  First and second order gradients as in the reference
(bias_act.py:142-203).  `impl` is accepted and ignored."""
import numpy as np
import torch

from ... import _lib


class _Spec(dict):
    __getattr__ = dict.__getitem__


# name -> defaults / native index / which forward tensors the gradient needs (reference bias_act.py:21-31)
activation_funcs = {
    'linear':   _Spec(def_alpha=0,   def_gain=1,          cuda_idx=1, ref='',  has_2nd_grad=False),
    'relu':     _Spec(def_alpha=0,   def_gain=np.sqrt(2), cuda_idx=2, ref='y', has_2nd_grad=False),
    'lrelu':    _Spec(def_alpha=0.2, def_gain=np.sqrt(2), cuda_idx=3, ref='y', has_2nd_grad=False),
    'tanh':     _Spec(def_alpha=0,   def_gain=1,          cuda_idx=4, ref='y', has_2nd_grad=True),
    'sigmoid':  _Spec(def_alpha=0,   def_gain=1,          cuda_idx=5, ref='y', has_2nd_grad=True),
    'elu':      _Spec(def_alpha=0,   def_gain=1,          cuda_idx=6, ref='y', has_2nd_grad=True),
    'selu':     _Spec(def_alpha=0,   def_gain=1,          cuda_idx=7, ref='y', has_2nd_grad=True),
    'softplus': _Spec(def_alpha=0,   def_gain=1,          cuda_idx=8, ref='y', has_2nd_grad=True),
    'swish':    _Spec(def_alpha=0,   def_gain=np.sqrt(2), cuda_idx=9, ref='x', has_2nd_grad=True),
}


def bias_act(x, b=None, dim=1, act='linear', alpha=None, gain=None, clamp=None, impl='cuda'):
    r"""y = clamp(act(x + b) * gain); arguments as in the reference (bias_act.py:61-80)."""
    assert isinstance(x, torch.Tensor)
    assert impl in ['ref', 'cuda']
    _lib.require_cuda(x, b)
    return _bias_act_cuda(dim=dim, act=act, alpha=alpha, gain=gain, clamp=clamp).apply(x, b)


def _native(x, b, xref, yref, dy, grad, dim, spec, alpha, gain, clamp):
    if x.dtype not in (torch.float32, torch.float16):
        raise RuntimeError('bias_act: x must be float16 or float32')
    assert x.is_contiguous()
    y = torch.empty_like(x)
    if x.numel() == 0:
        return y
    step = int(np.prod(x.shape[dim + 1:])) if b is not None else 1
    size = int(x.shape[dim]) if b is not None else 1
    for t in (b, xref, yref, dy):
        assert t is None or (t.is_contiguous() and t.dtype == x.dtype)
    _lib.check(_lib.lib().afcm_bias_act(_lib.ptr(x), _lib.ptr(b), _lib.ptr(xref), _lib.ptr(yref), _lib.ptr(dy),
                                        _lib.ptr(y), _lib.dtype_code(x.dtype), x.numel(), step, size, grad,
                                        spec.cuda_idx, alpha, gain, clamp, _lib.stream_ptr(x.device)))
    return y


_bias_act_cuda_cache = dict()


def _bias_act_cuda(dim=1, act='linear', alpha=None, gain=None, clamp=None):
    assert clamp is None or clamp >= 0
    spec = activation_funcs[act]
    alpha = float(alpha if alpha is not None else spec.def_alpha)
    gain = float(gain if gain is not None else spec.def_gain)
    clamp = float(clamp if clamp is not None else -1)
    key = (dim, act, alpha, gain, clamp)
    if key in _bias_act_cuda_cache:
        return _bias_act_cuda_cache[key]

    class BiasActCuda(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, b):
            x = x.contiguous()
            if b is not None:
                assert b.ndim == 1 and 0 <= dim < x.ndim and b.shape[0] == x.shape[dim]
                b = b.contiguous()
            y = x
            if act != 'linear' or gain != 1 or clamp >= 0 or b is not None:
                y = _native(x, b, None, None, None, 0, dim, spec, alpha, gain, clamp)
            need_x = 'x' in spec.ref or spec.has_2nd_grad
            # The reference plugin saves y only for activations whose derivative is expressed through it; for
            # 'linear' with a clamp it therefore never gates the gradient (bias_act.py:151-154), unlike its own
            # `_ref` path.  The `_ref` semantics are the oracle here, so y is also kept whenever a clamp is active.
            need_y = 'y' in spec.ref or (clamp >= 0 and 'x' not in spec.ref)
            ctx.save_for_backward(x if need_x else None, b if need_x else None, y if need_y else None)
            ctx.has_b = b is not None
            return y

        @staticmethod
        def backward(ctx, dy):
            dy = dy.contiguous()
            x, b, y = ctx.saved_tensors
            dx = None
            db = None
            if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
                dx = dy
                if act != 'linear' or gain != 1 or clamp >= 0:
                    dx = BiasActCudaGrad.apply(dy, x, b, y)
            if ctx.needs_input_grad[1] and ctx.has_b:
                db = dx.sum([i for i in range(dx.ndim) if i != dim])
            return dx, db

    class BiasActCudaGrad(torch.autograd.Function):
        @staticmethod
        def forward(ctx, dy, x, b, y):
            dx = _native(dy, b, x, y, None, 1, dim, spec, alpha, gain, clamp)
            ctx.save_for_backward(dy if spec.has_2nd_grad else None, x, b, y)
            return dx

        @staticmethod
        def backward(ctx, d_dx):
            d_dx = d_dx.contiguous()
            dy, x, b, y = ctx.saved_tensors
            d_dy = d_x = d_b = None
            if ctx.needs_input_grad[0]:
                d_dy = BiasActCudaGrad.apply(d_dx, x, b, y)
            if spec.has_2nd_grad and (ctx.needs_input_grad[1] or ctx.needs_input_grad[2]):
                d_x = _native(d_dx, b, x, y, dy, 2, dim, spec, alpha, gain, clamp)
            if spec.has_2nd_grad and ctx.needs_input_grad[2]:
                d_b = d_x.sum([i for i in range(d_x.ndim) if i != dim])
            return d_dy, d_x, d_b, None

    _bias_act_cuda_cache[key] = BiasActCuda
    return BiasActCuda
