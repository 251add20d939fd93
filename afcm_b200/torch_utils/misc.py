"""Host-side helpers with the semantics of the reference's torch_utils/misc.py that the generator
calls on its hot path (assert_shape misc.py:82, profiled_function :100, suppress_tracer_warnings :70)."""
import contextlib

import torch


def assert_shape(tensor, ref_shape):
    if tensor.ndim != len(ref_shape):
        raise AssertionError(f'Wrong number of dimensions: got {tensor.ndim}, expected {len(ref_shape)}')
    for idx, (size, ref_size) in enumerate(zip(tensor.shape, ref_shape)):
        if ref_size is None:
            continue
        if int(size) != int(ref_size):
            raise AssertionError(f'Wrong size for dimension {idx}: got {size}, expected {ref_size}')


def profiled_function(fn):
    def decorator(*args, **kwargs):
        with torch.autograd.profiler.record_function(fn.__name__):
            return fn(*args, **kwargs)
    decorator.__name__ = fn.__name__
    return decorator


@contextlib.contextmanager
def suppress_tracer_warnings():
    yield
