"""ctypes binding of libafcm_b200.so -- the only way Python reaches the CUDA kernels.

This is the binding a maintainer of the reference would add in place of
torch_utils/custom_ops.get_plugin (models/networks/stylegan3/torch_utils/custom_ops.py:59-155): no JIT
build at first call, no pybind/ATen types, just `extern "C"` entry points taking raw device pointers.
There is deliberately NO fallback: if the library is missing or an op is called on a non-CUDA tensor
the call raises.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# AFCM_B200_LIB: another build of the same library (A/B timing of kernel variants with tools/layer_bench.py)
LIB_PATH = os.environ.get('AFCM_B200_LIB') or os.path.join(_HERE, 'libafcm_b200.so')

OK, ERR_UNSUPPORTED, ERR_INVALID = 0, -1, -2
F32, F16, BF16 = 0, 1, 2
SIGN_NONE, SIGN_WRITE, SIGN_READ = 0, 1, 2

_c = ctypes
_vp, _i, _f, _i64 = _c.c_void_p, _c.c_int, _c.c_float, _c.c_int64
_pi = _c.POINTER(_c.c_int)

# name -> (restype, argtypes); mirrors include/afcm_b200.h one to one
SIGNATURES = {
    'afcm_version': (_i, []),
    'afcm_last_error': (_c.c_char_p, []),
    'afcm_device_check': (_i, []),
    'afcm_launch_count': (_c.c_longlong, []),
    'afcm_filtered_lrelu': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i,            # x xs y ys b skip dtype
                                 _i, _i, _i, _i, _i, _i,                       # N C xh xw yh yw
                                 _vp, _i, _vp, _i,                             # fu n fd n
                                 _i, _i, _i, _i, _i, _i,                       # up down px0 px1 py0 py1
                                 _f, _f, _f, _f, _i,                           # gain slope clamp out_scale flip
                                 _i, _vp, _i, _i, _i, _i, _vp]),               # sign_mode signs sh swb sx sy stream
    'afcm_filtered_lrelu_tcs': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _i, _vp, _i,
                                     _i, _i, _i, _i, _i, _i, _f, _f, _f, _f, _i, _i, _vp, _i, _i, _i, _i, _vp]),
    'afcm_filtered_lrelu_tc': (_i, [_vp, _vp, _i, _vp, _vp, _i, _vp, _vp,           # x xs xdt y ys ydt b skip
                                    _i, _i, _i, _i, _i, _i, _vp, _i, _vp, _i,        # N C xh xw yh yw fu n fd n
                                    _i, _i, _i, _i, _i, _i, _f, _f, _f, _f, _i, _vp]),
    'afcm_filtered_lrelu_tc_signs': (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _i, _vp, _i,
                                          _i, _i, _i, _i, _i, _i, _f, _f, _f, _i, _i, _vp, _i, _i, _i, _i, _vp, _vp]),
    'afcm_absmax': (_i, [_vp, _i64, _vp, _vp]),
    'afcm_filtered_lrelu_tc_padded': (_i, [_vp, _vp, _i, _vp, _vp, _i, _vp, _vp,           # x xs xdt y ys ydt b skip
                                    _i, _i, _i, _i, _i, _i, _vp, _i, _vp, _i,        # N C xh xw yh yw fu n fd n
                                    _i, _i, _i, _i, _i, _i, _f, _f, _f, _f, _i, _i, _vp]),  # up down pads gain slope clamp scale flip stream
    'afcm_filtered_lrelu_t5': (_i, [_vp, _vp, _i, _vp, _vp, _i, _vp, _vp,           # same arguments as afcm_filtered_lrelu_tc
                                    _i, _i, _i, _i, _i, _i, _vp, _i, _vp, _i,
                                    _i, _i, _i, _i, _i, _i, _f, _f, _f, _f, _i, _vp]),
    'afcm_filtered_lrelu_t5_plan': (_i, [_i] * 9 + [_pi, _i]),
    'afcm_filtered_lrelu_t5_trace': (_i, [_vp]),
    'afcm_filtered_lrelu_out_size': (_i, [_i] * 10 + [_pi, _pi]),
    'afcm_filtered_lrelu_sign_size': (_i, [_i] * 4 + [_pi, _pi]),
    'afcm_filtered_lrelu_set_tile': (_i, [_i, _i]),
    'afcm_filtered_lrelu_tc_set_waves': (_i, [_i]),
    'afcm_filtered_lrelu_act': (_i, [_vp, _i, _i64, _i, _i, _f, _f, _f, _i, _vp, _i, _i, _i, _i, _vp]),
    'afcm_upfirdn2d': (_i, [_vp, _vp, _i, _i64, _i, _i, _i, _i, _vp, _i, _i,
                            _i, _i, _i, _i, _i, _i, _i, _i, _i, _f, _vp]),
    'afcm_bias_act': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i64, _i64, _i64, _i, _i, _f, _f, _f, _vp]),
    'afcm_conv2d_f32': (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    'afcm_conv_weight_prep': (_i, [_vp, _i, _i, _i, _f, _i, _vp, _vp, _i, _vp, _vp]),
    'afcm_modconv_coefs': (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    'afcm_modconv_coefs_ema': (_i, [_vp, _vp, _vp, _i, _vp, _vp, _i, _i, _i, _i, _vp]),
    'afcm_conv_tc_plane_elems': (_i64, [_i, _i, _i]),
    'afcm_conv_tc_pack': (_i, [_vp, _i, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    'afcm_conv_tc_pack_pitched': (_i, [_vp, _i, _i, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    'afcm_conv2d_tc': (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    'afcm_conv2d_tc_nchw': (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    'afcm_plane_dot_scale': (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _vp]),
    'afcm_plane_sum': (_i, [_vp, _vp, _i64, _i64, _vp]),
    'afcm_conv2d_wgrad_f32': (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    'afcm_conv2d_wgrad_tc': (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    'afcm_conv2d_wgrad_tc_workspace': (_i64, [_i, _i, _i, _i, _i, _i]),
    'afcm_conv2d_wgrad_tc5': (_i, [_vp, _vp, _vp, _vp, _i64, _i, _i, _i, _i, _i, _i, _i, _vp]),
    'afcm_ema_lerp': (_i, [_vp, _vp, _i64, _f, _vp]),
    'afcm_adam_step': (_i, [_vp, _vp, _vp, _vp, _i64, _f, _f, _f, _f, _i, _f, _i, _vp]),
    'afcm_conv_tc_debug_buffer': (_vp, [_i]),
    'afcm_conv_tc_trace': (_i, [_vp]),
    'afcm_conv_tc_set_issuers': (_i, [_i]),
    'afcm_conv_tc_set_stages': (_i, [_i]),
    'afcm_conv_tc_set_rowreuse': (_i, [_i]),
    'afcm_fully_connected_grouped': (_i, [_i, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _i, _i, _f, _f, _vp]),
    'afcm_fully_connected': (_i, [_vp, _i64, _vp, _vp, _vp, _i64, _i, _i, _i, _f, _f, _i, _f, _f, _vp]),
    'afcm_normalize_2nd_moment': (_i, [_vp, _i64, _vp, _i64, _i, _i, _f, _vp]),
    'afcm_adaptive_avgpool': (_i, [_vp, _vp, _i64, _i, _i, _i, _i, _vp]),
    'afcm_pad_input': (_i, [_vp, _vp, _vp, _i64, _i, _i, _i, _vp]),
    'afcm_torgb': (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i64, _f, _f, _f, _f, _vp]),
    'afcm_fourier_features': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _f, _vp]),
}

_lib = None
_call_device = None          # device index of the tensors of the native call being assembled (set by stream_ptr())


class _DeviceGuarded:
    """The loaded library with a device guard around every entry point.  The C entry points launch on the calling thread's
    CURRENT device (cudaFuncSetAttribute, occupancy queries and the launch itself), while the stream handed to them belongs
    to the tensors' device: when the two differ the call is wrapped in `torch.cuda.device(...)` -- what OptionalCUDAGuard
    does in the reference plugins (OPS/filtered_lrelu.cpp:22).  The tensors' device is the one stream_ptr() was last asked
    for, which every wrapper evaluates while it assembles the arguments of its call."""

    def __init__(self, cdll):
        self._cdll = cdll
        self._fns = {}

    def __getattr__(self, name):
        fn = self._fns.get(name)
        if fn is None:
            raw = getattr(self._cdll, name)

            def fn(*args, _raw=raw):
                global _call_device
                dev, _call_device = _call_device, None
                if dev is not None:
                    import torch
                    if dev != torch.cuda.current_device():
                        with torch.cuda.device(dev):
                            return _raw(*args)
                return _raw(*args)
            self._fns[name] = fn
        return fn


def lib():
    """Loads the shared library (raises if it has not been built: `python -m afcm_b200.build`)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f'{LIB_PATH} not found: the CUDA extension is not built '
                               f'(run `python -m afcm_b200.build`). There is no CPU fallback.')
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            if not hasattr(L, name) and 'AFCM_B200_LIB' in os.environ:
                continue                                   # an older build lacks the newest tuning switches
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        if os.environ.get('AFCM_TC_ISSUERS') and hasattr(L, 'afcm_conv_tc_set_issuers'):       # tuning switch for A/B runs of bench.py
            L.afcm_conv_tc_set_issuers(int(os.environ['AFCM_TC_ISSUERS']))
        _lib = _DeviceGuarded(L)
    return _lib


def last_error():
    return lib().afcm_last_error().decode('utf-8', 'replace')


def check(rc, allow_unsupported=False):
    """0 -> ok; -1 -> returned to the caller when allow_unsupported (the reference's return_code < 0);
    anything else raises RuntimeError, the analogue of the reference's TORCH_CHECK / AT_CUDA_CHECK."""
    if rc == OK:
        return rc
    if rc == ERR_UNSUPPORTED and allow_unsupported:
        return rc
    raise RuntimeError(f'libafcm_b200 error {rc}: {last_error()}')


def launch_count():
    return int(lib().afcm_launch_count())


def dtype_code(t):
    import torch
    return {torch.float32: F32, torch.float16: F16, torch.bfloat16: BF16}[t]


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr(device=None):
    """Raw handle of torch's current stream on `device`; also notes the device for the guard around the native call."""
    import torch
    global _call_device
    if device is not None:
        device = torch.device(device)
        _call_device = device.index if device.index is not None else torch.cuda.current_device()
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError('afcm_b200 ops run on CUDA tensors only (sm_100a kernels; there is no CPU fallback)')


_host_cache = {}


def host_array(t, dtype=np.float32):
    """Host copy of a small constant tensor (filter taps), cached per tensor OBJECT (weak reference +
    version counter) so that the device->host sync happens once per filter, not per call.  Keying on the
    storage address alone would be wrong: the caching allocator hands freed addresses to new tensors."""
    import weakref
    if t is None:
        return None
    key = id(t)
    ent = _host_cache.get(key)
    if ent is not None and ent[0]() is t and ent[1] == t._version:
        return ent[2]
    a = np.ascontiguousarray(t.detach().to('cpu', copy=True).numpy().astype(dtype, copy=False))
    if len(_host_cache) > 1024:
        for k in [k for k, v in _host_cache.items() if v[0]() is None]:
            del _host_cache[k]
        if len(_host_cache) > 1024:
            _host_cache.clear()
    _host_cache[key] = (weakref.ref(t), t._version, a)
    return a


def np_ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def i64x4(vals):
    return (ctypes.c_int64 * 4)(*[int(v) for v in vals])


# ----------------------------------------------------------------------------------------------------
# Per-kernel timing used by bench.py's roofline leg: CUDA events on the launching stream around each
# heavy kernel launch.  Off (None) unless bench.py turns it on; the product path is unaffected.
_prof = None


def profile_begin():
    global _prof
    _prof = []


def profile_end():
    """-> {kernel name: dict(ms, work, launches)}; synchronises."""
    global _prof
    import torch
    torch.cuda.synchronize()
    rec, _prof = _prof, None
    out = {}
    for name, work, a, b in rec or []:
        d = out.setdefault(name, dict(ms=0.0, work=0.0, launches=0))
        d['ms'] += a.elapsed_time(b)
        d['work'] += work
        d['launches'] += 1
    return out


def timed(name, work, fn):
    """Runs fn(); when profiling is on, brackets it with CUDA events on the current stream."""
    if _prof is None:
        return fn()
    import torch
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    r = fn()
    b.record()
    _prof.append((name, float(work), a, b))
    return r
