"""Host-side mirror of the reference generator (models/networks/stylegan3/networks_stylegan3.py, cited
below as NET:line) on top of the afcm_b200 operator API.  Same class names, constructor arguments,
parameter / buffer names (so reference `*_net_G*.pth` state_dicts load unchanged)
and forward signatures; every tensor operation on the forward path is a hand-written sm_100a kernel
reached through libafcm_b200.so.  PyTorch provides parameters, device memory and streams only.

Every operator also carries its first-order autograd definition (the generator training step, BASELINE
config 5): native kernels for the convolution data / weight gradients, the filtered_lrelu backward (sign tensor),
bias_act and the FC products; tensor operations only on the small [N,C] / [O,I] coefficient tensors.
"""
import numpy as np
import scipy.signal
import scipy.special
import torch

from . import _lib
from .torch_utils import misc
from .torch_utils.ops import bias_act, conv2d_gradfix, filtered_lrelu
from .torch_utils.ops.filtered_lrelu import _run_fused as _flrelu_fused

# ----------------------------------------------------------------------------------------------------


@misc.profiled_function
def modulated_conv2d(x, w, s, demodulate=True, padding=0, input_gain=None, impl=None, out_dtype=None, bias=None, input_gain_rsqrt=False):
    """NET:25-64.  x [N,I,H,W], w [O,I,k,k], s [N,I]; input_gain [] / [I] / [N,I] or None.

    The reference materialises a per-sample weight w*s*d*g and runs a grouped conv.  Here the same
    product is evaluated as  d[n,o] * conv(x * (s_hat*g)[n,i], w_hat)  -- modulation on the activation
    side, demodulation in the GEMM epilogue -- so no [N,O,I,k,k] tensor exists.
    `input_gain_rsqrt` (not in the reference signature): `input_gain` is the layer's magnitude_ema buffer itself and the
    coefficient kernel applies the rsqrt of NET:346.
    """
    _lib.require_cuda(x, w, s)
    N = int(x.shape[0])
    O, I, kh, kw = w.shape
    misc.assert_shape(w, [O, I, kh, kw])
    misc.assert_shape(x, [N, I, None, None])
    misc.assert_shape(s, [N, I])
    if conv2d_gradfix.needs_grad(x, w, s):
        return _modulated_conv2d_train(x, w, s, demodulate, padding, input_gain.rsqrt() if input_gain_rsqrt else input_gain, impl, bias)
    L = _lib.lib()
    s = s.contiguous().float()
    gain_scalar = None
    if input_gain is not None:
        if input_gain.numel() == 1:
            gain_scalar = input_gain.reshape(1).float().contiguous()
        else:
            input_gain = input_gain.rsqrt() if input_gain_rsqrt else input_gain
            input_gain_rsqrt = False
            s = (s * input_gain.expand(N, I)) if not demodulate else s    # per-channel gains handled below
    ent = conv2d_gradfix.prepare_weight(w, 1.0, bool(demodulate), want_wsq=bool(demodulate))
    icoef = torch.empty([N, I], dtype=torch.float32, device=x.device)
    ocoef = torch.empty([N, O], dtype=torch.float32, device=x.device) if demodulate else None
    _lib.check(L.afcm_modconv_coefs_ema(_lib.ptr(s), _lib.ptr(ent.get('wsq')), _lib.ptr(gain_scalar), int(bool(input_gain_rsqrt)),
                                        _lib.ptr(icoef), _lib.ptr(ocoef), N, I, O, int(bool(demodulate)), _lib.stream_ptr(x.device)))
    if input_gain is not None and input_gain.numel() != 1 and demodulate:
        icoef = icoef * input_gain.expand(N, I)        # rare general form (NET:55-57); AFCM passes a scalar
    return conv2d_gradfix.conv2d_native(x, w, int(padding), icoef=icoef, ocoef=ocoef, pre_scale=1.0,
                                        normalize=bool(demodulate), impl=impl, out_dtype=out_dtype, bias=bias)


def _modulated_conv2d_train(x, w, s, demodulate, padding, input_gain, impl, bias):
    """Differentiable modulated_conv2d (training step).  The coefficient algebra of NET:41-57 runs as tensor operations
    on the small [N,I] / [O,I] tensors so that autograd carries it; the demodulation GEMV and the convolution with its
    three gradients are native kernels (_FcLinearFn, conv2d_gradfix._ConvFn)."""
    N, I = s.shape
    s = s.float()
    wp = w.float()
    ocoef = None
    if demodulate:
        wp = wp * wp.square().mean([1, 2, 3], keepdim=True).rsqrt()                  # NET:42
        s = s * s.square().mean().rsqrt()                                            # NET:43
        wsq = wp.square().sum([2, 3])                                                # [O, I]
        ocoef = (_FcLinearFn.apply(s.square(), wsq, 1.0) + 1e-8).rsqrt()             # NET:51  [N, O]
    icoef = s if input_gain is None else s * input_gain.expand(N, I)                 # NET:55-57
    y = conv2d_gradfix._ConvFn.apply(x, wp, icoef, ocoef, int(padding), impl or conv2d_gradfix.conv_impl)
    if bias is not None:
        y = y + bias.reshape(1, -1, 1, 1)
    return y


# ----------------------------------------------------------------------------------------------------


class _FcLinearFn(torch.autograd.Function):
    """y = x @ (w * weight_gain)^T with the native FC kernel for the product and for both gradients
    (dx = dy @ w, dw = dy^T @ x): the matmul / addmm of NET:97-99 and its autograd."""

    @staticmethod
    def _mm(a, b, gain):                     # a [M,K], b [P,K] -> a @ b^T * gain
        a = a.contiguous().float()
        b = b.contiguous().float()
        M, K = a.shape
        P = b.shape[0]
        y = torch.empty([M, P], dtype=torch.float32, device=a.device)
        _lib.check(_lib.lib().afcm_fully_connected(_lib.ptr(a), a.stride(0), _lib.ptr(b), None, _lib.ptr(y), y.stride(0),
                                                   M, K, P, float(gain), 1.0, 1, 0.0, 1.0, _lib.stream_ptr(a.device)))
        return y

    @staticmethod
    def forward(ctx, x, w, weight_gain):
        _lib.require_cuda(x, w)
        ctx.save_for_backward(x, w)
        ctx.gain = float(weight_gain)
        return _FcLinearFn._mm(x, w, ctx.gain)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = _FcLinearFn._mm(dy, w.t(), ctx.gain)
        if ctx.needs_input_grad[1]:
            dw = _FcLinearFn._mm(dy.t(), x.t(), ctx.gain)
        return dx, dw, None


def _fc_train(x, weight, bias, weight_gain, bias_gain, activation):
    """FullyConnectedLayer.forward with an autograd graph, structured like NET:89-101: product, then bias_act."""
    y = _FcLinearFn.apply(x, weight, weight_gain)
    b = bias
    if b is not None and bias_gain != 1:
        b = b * bias_gain
    if activation == 'linear':
        return y if b is None else y + b.unsqueeze(0)
    return bias_act.bias_act(y, b, act=activation)


def _fc_native(x, weight, bias, weight_gain, bias_gain, activation, out=None):
    spec = bias_act.activation_funcs[activation]
    _lib.require_cuda(x, weight)
    if out is None and conv2d_gradfix.needs_grad(x, weight, bias):
        return _fc_train(x, weight, bias, weight_gain, bias_gain, activation)
    assert x.ndim == 2 and x.dtype == torch.float32 and x.stride(1) == 1
    N, in_f = x.shape
    out_f = weight.shape[0]
    y = out if out is not None else torch.empty([N, out_f], dtype=torch.float32, device=x.device)
    assert y.stride(1) == 1
    w = weight.detach().contiguous()
    b = bias.detach().contiguous() if bias is not None else None
    act_gain = float(spec.def_gain)
    _lib.check(_lib.lib().afcm_fully_connected(
        _lib.ptr(x), x.stride(0), _lib.ptr(w), _lib.ptr(b), _lib.ptr(y), y.stride(0), N, in_f, out_f,
        float(weight_gain), float(bias_gain), spec.cuda_idx, float(spec.def_alpha), act_gain, _lib.stream_ptr(x.device)))
    return y


class FullyConnectedLayer(torch.nn.Module):
    """NET:69-101."""

    def __init__(self, in_features, out_features, activation='linear', bias=True, lr_multiplier=1, weight_init=1,
                 bias_init=0):
        super().__init__()
        self.in_features = in_features
        self.out_features = out_features
        self.activation = activation
        self.weight = torch.nn.Parameter(torch.randn([out_features, in_features]) * (weight_init / lr_multiplier))
        bias_init = np.broadcast_to(np.asarray(bias_init, dtype=np.float32), [out_features])
        self.bias = torch.nn.Parameter(torch.from_numpy(bias_init / lr_multiplier)) if bias else None
        self.weight_gain = lr_multiplier / np.sqrt(in_features)
        self.bias_gain = lr_multiplier

    def forward(self, x, out=None):
        return _fc_native(x.float(), self.weight, self.bias, self.weight_gain, self.bias_gain, self.activation, out=out)

    def extra_repr(self):
        return f'in_features={self.in_features:d}, out_features={self.out_features:d}, activation={self.activation:s}'


class MappingNetwork(torch.nn.Module):
    """NET:109-161."""

    def __init__(self, z_dim, c_dim, w_dim, num_ws, num_layers=2, lr_multiplier=0.01, w_avg_beta=0.998):
        super().__init__()
        self.z_dim, self.c_dim, self.w_dim, self.num_ws = z_dim, c_dim, w_dim, num_ws
        self.num_layers = num_layers
        self.w_avg_beta = w_avg_beta
        self.embed = FullyConnectedLayer(c_dim, w_dim) if c_dim > 0 else None
        features = [z_dim + (w_dim if c_dim > 0 else 0)] + [w_dim] * num_layers
        for idx, in_f, out_f in zip(range(num_layers), features[:-1], features[1:]):
            setattr(self, f'fc{idx}', FullyConnectedLayer(in_f, out_f, activation='lrelu', lr_multiplier=lr_multiplier))
        self.register_buffer('w_avg', torch.zeros([w_dim]))

    def forward(self, z, c, truncation_psi=1, truncation_cutoff=None, update_emas=False, **kwargs):
        misc.assert_shape(z, [None, self.z_dim])
        _lib.require_cuda(z)
        if truncation_cutoff is None:
            truncation_cutoff = self.num_ws
        L = _lib.lib()
        N = z.shape[0]
        st = _lib.stream_ptr(z.device)
        width = self.z_dim + (self.w_dim if self.c_dim > 0 else 0)
        if conv2d_gradfix.needs_grad(z, c, *self.parameters()):
            return self._forward_train(z, c, truncation_psi, truncation_cutoff, update_emas)
        x = torch.empty([N, width], dtype=torch.float32, device=z.device)
        z = z.float().contiguous()
        _lib.check(L.afcm_normalize_2nd_moment(_lib.ptr(z), z.stride(0), _lib.ptr(x), x.stride(0), N, self.z_dim, 1e-8, st))
        if self.c_dim > 0:
            misc.assert_shape(c, [None, self.c_dim])
            y = self.embed(c.float().contiguous())
            xv = x[:, self.z_dim:]
            _lib.check(L.afcm_normalize_2nd_moment(_lib.ptr(y), y.stride(0), _lib.ptr(xv), x.stride(0), N, self.w_dim, 1e-8, st))
        for idx in range(self.num_layers):
            x = getattr(self, f'fc{idx}')(x)
        if update_emas:
            self.w_avg.copy_(x.detach().mean(dim=0).lerp(self.w_avg, self.w_avg_beta))
        x = x.unsqueeze(1).repeat([1, self.num_ws, 1])
        if truncation_psi != 1:
            x[:, :truncation_cutoff] = self.w_avg.lerp(x[:, :truncation_cutoff], truncation_psi)
        return x

    def _forward_train(self, z, c, truncation_psi, truncation_cutoff, update_emas):
        """NET:141-158 with an autograd graph: the FC products and activations are native (_fc_train), the 2nd-moment
        normalisation of the two [N,512] codes is written with tensor operations."""
        def norm2(v):
            return v * (v.square().mean(dim=1, keepdim=True) + 1e-8).rsqrt()
        x = norm2(z.float())
        if self.c_dim > 0:
            x = torch.cat([x, norm2(self.embed(c.float().contiguous()))], dim=1)
        for idx in range(self.num_layers):
            x = getattr(self, f'fc{idx}')(x)
        if update_emas:
            self.w_avg.copy_(x.detach().mean(dim=0).lerp(self.w_avg, self.w_avg_beta))
        x = x.unsqueeze(1).repeat([1, self.num_ws, 1])
        if truncation_psi != 1:
            x[:, :truncation_cutoff] = self.w_avg.lerp(x[:, :truncation_cutoff], truncation_psi)
        return x

    def extra_repr(self):
        return f'z_dim={self.z_dim:d}, c_dim={self.c_dim:d}, w_dim={self.w_dim:d}, num_ws={self.num_ws:d}'


class SynthesisInput(torch.nn.Module):
    """NET:169-243 (Fourier features).  Part of the operator surface named by the north star; the AFCM
    network itself never instantiates it (NET:641-645)."""

    def __init__(self, w_dim, channels, size, sampling_rate, bandwidth):
        super().__init__()
        self.w_dim = w_dim
        self.channels = channels
        self.size = np.broadcast_to(np.asarray(size), [2])
        self.sampling_rate = sampling_rate
        self.bandwidth = bandwidth
        freqs = torch.randn([self.channels, 2])
        radii = freqs.square().sum(dim=1, keepdim=True).sqrt()
        freqs /= radii * radii.square().exp().pow(0.25)
        freqs *= bandwidth
        phases = torch.rand([self.channels]) - 0.5
        self.weight = torch.nn.Parameter(torch.randn([self.channels, self.channels]))
        self.affine = FullyConnectedLayer(w_dim, 4, weight_init=0, bias_init=[1, 0, 0, 0])
        self.register_buffer('transform', torch.eye(3, 3))
        self.register_buffer('freqs', freqs)
        self.register_buffer('phases', phases)

    def forward(self, w):
        t = self.affine(w)
        N = w.shape[0]
        H, W = int(self.size[1]), int(self.size[0])
        y = torch.empty([N, self.channels, H, W], dtype=torch.float32, device=w.device)
        _lib.check(_lib.lib().afcm_fourier_features(
            _lib.ptr(t), _lib.ptr(self.freqs.contiguous()), _lib.ptr(self.phases.contiguous()),
            _lib.ptr(self.weight.detach().contiguous()), _lib.ptr(self.transform.contiguous()), _lib.ptr(y),
            N, self.channels, H, W, float(self.sampling_rate), float(self.bandwidth), _lib.stream_ptr(w.device)))
        return y


# ----------------------------------------------------------------------------------------------------


def design_lowpass_filter(numtaps, cutoff, width, fs, radial=False):
    """NET:381-402 / NET:518-539: Kaiser low-pass via scipy.signal.firwin (None = identity)."""
    assert numtaps >= 1
    if numtaps == 1:
        return None
    if not radial:
        f = scipy.signal.firwin(numtaps=numtaps, cutoff=cutoff, width=width, fs=fs)
        return torch.as_tensor(f, dtype=torch.float32)
    x = (np.arange(numtaps) - (numtaps - 1) / 2) / fs
    r = np.hypot(*np.meshgrid(x, x))
    f = scipy.special.j1(2 * cutoff * (np.pi * r)) / (np.pi * r)
    beta = scipy.signal.kaiser_beta(scipy.signal.kaiser_atten(numtaps, width / (fs / 2)))
    w = np.kaiser(numtaps, beta)
    f *= np.outer(w, w)
    f /= np.sum(f)
    return torch.as_tensor(f, dtype=torch.float32)


class _AliasFreeLayerBase(torch.nn.Module):
    """Filter design and padding shared by SynthesisLayer (NET:294-334) and EncoderLayer (NET:453-489)."""

    def _setup_filters(self, in_size, out_size, in_sampling_rate, out_sampling_rate, in_cutoff, out_cutoff,
                       in_half_width, out_half_width, conv_kernel, filter_size, lrelu_upsampling, use_radial_filters,
                       is_torgb, is_critically_sampled):
        self.in_size = np.broadcast_to(np.asarray(in_size), [2])
        self.out_size = np.broadcast_to(np.asarray(out_size), [2])
        self.in_sampling_rate = in_sampling_rate
        self.out_sampling_rate = out_sampling_rate
        self.tmp_sampling_rate = max(in_sampling_rate, out_sampling_rate) * (1 if is_torgb else lrelu_upsampling)
        self.in_cutoff, self.out_cutoff = in_cutoff, out_cutoff
        self.in_half_width, self.out_half_width = in_half_width, out_half_width
        self.conv_kernel = 1 if is_torgb else conv_kernel

        self.up_factor = int(np.rint(self.tmp_sampling_rate / self.in_sampling_rate))
        assert self.in_sampling_rate * self.up_factor == self.tmp_sampling_rate
        self.up_taps = filter_size * self.up_factor if self.up_factor > 1 and not is_torgb else 1
        self.register_buffer('up_filter', design_lowpass_filter(
            numtaps=self.up_taps, cutoff=self.in_cutoff, width=self.in_half_width * 2, fs=self.tmp_sampling_rate))

        self.down_factor = int(np.rint(self.tmp_sampling_rate / self.out_sampling_rate))
        assert self.out_sampling_rate * self.down_factor == self.tmp_sampling_rate
        self.down_taps = filter_size * self.down_factor if self.down_factor > 1 and not is_torgb else 1
        self.down_radial = use_radial_filters and not is_critically_sampled
        self.register_buffer('down_filter', design_lowpass_filter(
            numtaps=self.down_taps, cutoff=self.out_cutoff, width=self.out_half_width * 2, fs=self.tmp_sampling_rate,
            radial=self.down_radial))

        pad_total = (self.out_size - 1) * self.down_factor + 1
        pad_total -= (self.in_size + self.conv_kernel - 1) * self.up_factor
        pad_total += self.up_taps + self.down_taps - 2
        pad_lo = (pad_total + self.up_factor) // 2
        pad_hi = pad_total - pad_lo
        self.padding = [int(pad_lo[0]), int(pad_hi[0]), int(pad_lo[1]), int(pad_hi[1])]

    def _filtered_lrelu(self, x, gain, slope, skip=None, out_scale=1.0, out_dtype=None, bias_done=False, conv_ready=False):
        """bias + filtered leaky ReLU + clamp (NET:371-372 / NET:510-511), with the skip addition
        (NET:376-377) and the output scale (NET:699-700) folded into the kernel epilogue when no autograd
        graph is needed.  On the fast inference path (conv2d_gradfix.fast_path()) the tensor-core kernel
        runs it on fp16 planes; `out_dtype` then selects the storage type of the result.  `conv_ready`: the result feeds a
        3x3 convolution -- it is stored at the row pitch W + 2 with two zero columns behind every row, the flat plane the
        tcgen05 GEMM reads directly (no pack pass between the two operators)."""
        b = None if bias_done else self.bias.to(torch.float32)    # bias_done: the conv epilogue already added it
        needs_graph = torch.is_grad_enabled() and (x.requires_grad or (b is not None and b.requires_grad))
        if not needs_graph:
            px0, px1, py0, py1 = self.padding
            clamp = float(self.conv_clamp) if self.conv_clamp is not None else float('inf')
            if conv2d_gradfix.fast_path() and self.up_filter is not None and self.down_filter is not None:
                y = filtered_lrelu.filtered_lrelu_tc(x, self.up_filter, self.down_filter, b, up=self.up_factor,
                                                     down=self.down_factor, padding=self.padding, gain=float(gain),
                                                     slope=float(slope), clamp=self.conv_clamp, skip=skip,
                                                     out_scale=float(out_scale), out_dtype=out_dtype or x.dtype,
                                                     conv_ready=conv_ready)
                if y is not None:
                    return y
            if x.dtype == torch.float32:
                y, _, rc = _flrelu_fused(x, self.up_filter, self.down_filter, b, None, self.up_factor, self.down_factor,
                                         px0, px1, py0, py1, 0, 0, float(gain), float(slope), clamp, False, False,
                                         skip=skip, out_scale=float(out_scale))
                if rc == 0:
                    return y
        y = filtered_lrelu.filtered_lrelu(x=x, fu=self.up_filter, fd=self.down_filter, b=None if b is None else b.to(x.dtype), up=self.up_factor,
                                          down=self.down_factor, padding=self.padding, gain=gain, slope=slope,
                                          clamp=self.conv_clamp)
        if skip is not None:
            y = y + skip
        return y * out_scale if out_scale != 1.0 else y


class SynthesisLayer(_AliasFreeLayerBase):
    """NET:253-379."""

    def __init__(self, w_dim, global_w_dim, is_torgb, is_critically_sampled, use_fp16, in_channels, out_channels,
                 in_size, out_size, in_sampling_rate, out_sampling_rate, in_cutoff, out_cutoff, in_half_width,
                 out_half_width, conv_kernel=3, filter_size=6, lrelu_upsampling=2, use_radial_filters=False,
                 conv_clamp=256, magnitude_ema_beta=0.999, cond_mod=False):
        super().__init__()
        self.w_dim = w_dim
        self.is_torgb = is_torgb
        self.is_critically_sampled = is_critically_sampled
        self.use_fp16 = use_fp16
        self.in_channels, self.out_channels = in_channels, out_channels
        self.conv_clamp = conv_clamp
        self.magnitude_ema_beta = magnitude_ema_beta
        self.cond_mod = cond_mod
        if not cond_mod:
            global_w_dim = 0
        k = 1 if is_torgb else conv_kernel
        self.affine = FullyConnectedLayer(self.w_dim + global_w_dim, self.in_channels, bias_init=1)
        self.weight = torch.nn.Parameter(torch.randn([self.out_channels, self.in_channels, k, k]))
        self.bias = torch.nn.Parameter(torch.zeros([self.out_channels]))
        self.register_buffer('magnitude_ema', torch.ones([]))
        self._setup_filters(in_size, out_size, in_sampling_rate, out_sampling_rate, in_cutoff, out_cutoff, in_half_width,
                            out_half_width, conv_kernel, filter_size, lrelu_upsampling, use_radial_filters, is_torgb,
                            is_critically_sampled)

    def _torgb_fused(self, x, styles, input_gain, out_scale, gain_rsqrt=False):
        """ToRGB in one kernel (afcm_torgb): 1x1 modulated conv without demodulation + bias + clamp + output scale."""
        L = _lib.lib()
        N, I, H, W = x.shape
        O = self.out_channels
        if x.dtype not in (torch.float32, torch.float16) or not x.is_contiguous() or input_gain.numel() != 1:
            return None
        s = styles.contiguous().float()
        icoef = torch.empty([N, I], dtype=torch.float32, device=x.device)
        st = _lib.stream_ptr(x.device)
        _lib.check(L.afcm_modconv_coefs_ema(_lib.ptr(s), None, _lib.ptr(input_gain.reshape(1).float().contiguous()), int(bool(gain_rsqrt)),
                                            _lib.ptr(icoef), None, N, I, O, 0, st))
        y = torch.empty([N, O, H, W], dtype=torch.float32, device=x.device)
        clamp = float(self.conv_clamp) if self.conv_clamp is not None else -1.0
        rc = _lib.timed('torgb', float(x.element_size() * x.numel() + 4 * y.numel()), lambda: L.afcm_torgb(
            _lib.ptr(x), _lib.dtype_code(x.dtype), _lib.ptr(self.weight.detach().reshape(O, I).float().contiguous()),
            _lib.ptr(icoef), None, _lib.ptr(self.bias.detach().float().contiguous()), _lib.ptr(y), N, I, O, H * W,
            1.0, 1.0, clamp, float(out_scale), st))
        _lib.check(rc, allow_unsupported=True)
        return y if rc == 0 else None

    def forward(self, x, w, global_w, E_features=None, include_skip=True, noise_mode='random', force_fp32=False,
                update_emas=False, out_scale=1.0, out_dtype=None, conv_ready=False, styles=None):
        assert noise_mode in ['random', 'const', 'none']
        misc.assert_shape(x, [None, self.in_channels, int(self.in_size[1]), int(self.in_size[0])])
        # `styles`: this layer's affine output, already computed by the caller (SynthesisNetwork runs the affine layers of all
        # synthesis layers in one grouped launch, incl. the ToRGB scale of NET:353-355)
        # `global_w is None`: the caller hands over the concatenation [w, img_global] itself (SynthesisNetwork builds it for all
        # layers with one kernel instead of one torch.cat per layer, NET:349-352)
        misc.assert_shape(w, [x.shape[0], self.w_dim + (self.affine.in_features - self.w_dim if global_w is None else 0)])
        if update_emas:
            magnitude_cur = x.detach().to(torch.float32).square().mean()
            self.magnitude_ema.copy_(magnitude_cur.lerp(self.magnitude_ema, self.magnitude_ema_beta))
        # no autograd graph: the coefficient kernels apply rsqrt(magnitude_ema) themselves (NET:346), no tiny kernel per layer
        ema_in_kernel = not torch.is_grad_enabled()
        input_gain = self.magnitude_ema if ema_in_kernel else self.magnitude_ema.rsqrt()
        if styles is None:
            if self.cond_mod and global_w is not None:
                w = torch.cat((w, global_w), 1)
            styles = self.affine(w)
            if self.is_torgb:
                styles = styles * (1 / np.sqrt(self.in_channels * (self.conv_kernel ** 2)))
        x_skip = None
        if E_features is not None and include_skip:
            x_skip = E_features[self.out_size[0]]
        fast = conv2d_gradfix.fast_path() and not torch.is_grad_enabled()
        if fast and self.is_torgb:
            y = self._torgb_fused(x, styles, input_gain, out_scale, ema_in_kernel)
            if y is not None:
                return y
        fast = fast and not self.is_torgb
        x = modulated_conv2d(x=x if fast else x.float(), w=self.weight, s=styles, padding=self.conv_kernel - 1,
                             demodulate=(not self.is_torgb), input_gain=input_gain, input_gain_rsqrt=ema_in_kernel,
                             out_dtype=conv2d_gradfix.act_dtype if fast else None,
                             bias=self.bias.detach().float() if fast else None)
        gain = 1 if self.is_torgb else np.sqrt(2)
        slope = 1 if self.is_torgb else 0.2
        x = self._filtered_lrelu(x, gain, slope, skip=x_skip if include_skip else None, out_scale=out_scale,
                                 out_dtype=out_dtype, bias_done=fast, conv_ready=conv_ready and fast)
        misc.assert_shape(x, [None, self.out_channels, int(self.out_size[1]), int(self.out_size[0])])
        return x


class EncoderLayer(_AliasFreeLayerBase):
    """NET:417-516."""

    def __init__(self, is_critically_sampled, use_fp16, in_channels, out_channels, in_size, out_size,
                 in_sampling_rate, out_sampling_rate, in_cutoff, out_cutoff, in_half_width, out_half_width,
                 conv_kernel=3, filter_size=6, lrelu_upsampling=1, use_radial_filters=False, conv_clamp=256,
                 magnitude_ema_beta=0.999, cond_mod=False):
        super().__init__()
        self.is_critically_sampled = is_critically_sampled
        self.use_fp16 = use_fp16
        self.in_channels, self.out_channels = in_channels, out_channels
        self.conv_clamp = conv_clamp
        self.magnitude_ema_beta = magnitude_ema_beta
        self.weight = torch.nn.Parameter(torch.randn([self.out_channels, self.in_channels, conv_kernel, conv_kernel]))
        self.weight_gain = 1 / np.sqrt(in_channels * (conv_kernel ** 2))
        self.bias = torch.nn.Parameter(torch.zeros([self.out_channels]))
        self.register_buffer('magnitude_ema', torch.ones([]))
        self._setup_filters(in_size, out_size, in_sampling_rate, out_sampling_rate, in_cutoff, out_cutoff, in_half_width,
                            out_half_width, conv_kernel, filter_size, lrelu_upsampling, use_radial_filters, False,
                            is_critically_sampled)

    def forward(self, x, force_fp32=False, update_emas=False, conv_ready=False):
        misc.assert_shape(x, [None, self.in_channels, int(self.in_size[1]), int(self.in_size[0])])
        if update_emas:
            magnitude_cur = x.detach().to(torch.float32).square().mean()
            self.magnitude_ema.copy_(magnitude_cur.lerp(self.magnitude_ema, self.magnitude_ema_beta))
        fast = conv2d_gradfix.fast_path() and not torch.is_grad_enabled()
        if conv2d_gradfix.needs_grad(x, self.weight):
            x = conv2d_gradfix.conv2d_train(x, self.weight, self.conv_kernel - 1, pre_scale=self.weight_gain)
        else:
            x = conv2d_gradfix.conv2d_native(x if fast else x.float(), self.weight, self.conv_kernel - 1, pre_scale=self.weight_gain,
                                             out_dtype=conv2d_gradfix.act_dtype if fast else None,
                                             bias=self.bias.detach().float() if fast else None)
        x = self._filtered_lrelu(x, np.sqrt(2), 0.2, bias_done=fast, conv_ready=conv_ready and fast)
        misc.assert_shape(x, [None, self.out_channels, int(self.out_size[1]), int(self.out_size[0])])
        return x


class Conv2dLayer(torch.nn.Module):
    """models/networks/CoModGAN/layers.py:116-162, restricted to what the AFCM generator instantiates
    (e_16x16: up = down = 1).  conv + bias_act."""

    def __init__(self, in_channels, out_channels, kernel_size, bias=True, activation='linear', up=1, down=1,
                 resample_filter=[1, 3, 3, 1], conv_clamp=None, channels_last=False, trainable=True):
        super().__init__()
        if up != 1 or down != 1:
            raise NotImplementedError('Conv2dLayer with resampling is not on the AFCM stylegan3 generator path')
        from .torch_utils.ops import upfirdn2d
        self.activation = activation
        self.up, self.down = up, down
        self.conv_clamp = conv_clamp
        self.register_buffer('resample_filter', upfirdn2d.setup_filter(resample_filter))
        self.padding = kernel_size // 2
        self.weight_gain = 1 / np.sqrt(in_channels * (kernel_size ** 2))
        self.act_gain = bias_act.activation_funcs[activation].def_gain
        weight = torch.randn([out_channels, in_channels, kernel_size, kernel_size])
        bias = torch.zeros([out_channels]) if bias else None
        if trainable:
            self.weight = torch.nn.Parameter(weight)
            self.bias = torch.nn.Parameter(bias) if bias is not None else None
        else:
            self.register_buffer('weight', weight)
            if bias is not None:
                self.register_buffer('bias', bias)
            else:
                self.bias = None

    def forward(self, x, gain=1):
        if conv2d_gradfix.needs_grad(x, self.weight):
            x = conv2d_gradfix.conv2d_train(x, self.weight, self.padding, pre_scale=self.weight_gain)
        else:
            x = conv2d_gradfix.conv2d_native(x if conv2d_gradfix.fast_path() else x.float(), self.weight, self.padding,
                                             pre_scale=self.weight_gain)
        b = self.bias.to(x.dtype) if self.bias is not None else None
        act_gain = self.act_gain * gain
        act_clamp = self.conv_clamp * gain if self.conv_clamp is not None else None
        return bias_act.bias_act(x, b, act=self.activation, gain=act_gain, clamp=act_clamp)


class _AvgPool4Fn(torch.autograd.Function):
    """AdaptiveAvgPool2d((4,4)) of NET:636,683: native forward; the gradient of a uniform window average is the
    output gradient spread over each window (sizes divisible by 4, which is what the generator produces)."""

    @staticmethod
    def forward(ctx, x):
        N, C, H, W = x.shape
        y = torch.empty([N, C, 4, 4], dtype=torch.float32, device=x.device)
        _lib.check(_lib.lib().afcm_adaptive_avgpool(_lib.ptr(x.contiguous().float()), _lib.ptr(y), N * C, H, W, 4, 4,
                                                    _lib.stream_ptr(x.device)))
        ctx.hw = (H, W)
        return y

    @staticmethod
    def backward(ctx, dy):
        H, W = ctx.hw
        if H % 4 or W % 4:
            raise NotImplementedError('adaptive average pool gradient: plane size must be divisible by 4')
        kh, kw = H // 4, W // 4
        return (dy * (1.0 / (kh * kw))).repeat_interleave(kh, dim=2).repeat_interleave(kw, dim=3)


class SynthesisNetwork(torch.nn.Module):
    """NET:556-705: 14 alias-free encoder layers, pooled global code, 15 co-modulated synthesis layers."""

    def __init__(self, w_dim, img_resolution, img_channels_in, img_channels_out, channel_base=32768, channel_max=512,
                 num_layers=14, num_critical=2, first_cutoff=2, first_stopband=2 ** 2.1, last_stopband_rel=2 ** 0.3,
                 margin_size=10, output_scale=0.25, num_fp16_res=4, dropout_rate=0.5, skip_resolution=256,
                 **layer_kwargs):
        super().__init__()
        self.w_dim = w_dim
        self.num_ws = num_layers + 2
        self.img_resolution = img_resolution
        self.img_channels_in, self.img_channels_out = img_channels_in, img_channels_out
        self.num_layers, self.num_critical = num_layers, num_critical
        self.margin_size = margin_size
        self.output_scale = output_scale
        self.num_fp16_res = num_fp16_res
        self.img_resolution_log2 = int(np.log2(img_resolution))
        if skip_resolution >= 4:
            final_skip = int(np.log2(skip_resolution))
            self.skip_connects = [True] * (final_skip - 1) + [False] * (self.img_resolution_log2 - final_skip)
        else:
            self.skip_connects = [False] * self.img_resolution_log2

        # geometric progression of cutoffs / stopbands (NET:595-611)
        last_cutoff = self.img_resolution / 2
        last_stopband = last_cutoff * last_stopband_rel
        exponents = np.minimum(np.arange(self.num_layers + 1) / (self.num_layers - self.num_critical), 1)
        cutoffs = first_cutoff * (last_cutoff / first_cutoff) ** exponents
        stopbands = first_stopband * (last_stopband / first_stopband) ** exponents
        sampling_rates = np.exp2(np.ceil(np.log2(np.minimum(stopbands * 2, self.img_resolution))))
        half_widths = np.maximum(stopbands, sampling_rates / 2) - cutoffs
        sizes = sampling_rates + self.margin_size * 2
        sizes_for_encoder = sizes.copy()
        sizes[-2:] = self.img_resolution
        self.sizes = sizes
        self.channels = channels = np.rint(np.minimum((channel_base / 2) / cutoffs, channel_max))
        channels[-1] = self.img_channels_out

        for idx in range(self.num_layers):
            rev_idx = self.num_layers - idx - 1
            rev_prev = self.num_layers - max(idx - 1, 0) - 1
            layer = EncoderLayer(
                is_critically_sampled=(idx < self.num_layers - self.num_critical), use_fp16=False,
                in_channels=self.img_channels_in if idx == 0 else int(channels[rev_prev]),
                out_channels=int(channels[rev_idx]),
                in_size=int(sizes_for_encoder[rev_prev]), out_size=int(sizes_for_encoder[rev_idx]),
                in_sampling_rate=int(sampling_rates[rev_prev]), out_sampling_rate=int(sampling_rates[rev_idx]),
                in_cutoff=cutoffs[rev_prev], out_cutoff=cutoffs[rev_idx],
                in_half_width=half_widths[rev_prev], out_half_width=half_widths[rev_idx], **layer_kwargs)
            setattr(self, f'encoder_{idx}', layer)

        self.e_16x16 = Conv2dLayer(int(channels[0]), int(channels[0]), kernel_size=3, activation='lrelu', conv_clamp=None)
        self.fc_in = FullyConnectedLayer(int(channels[0]) * (4 ** 2), 512 * 2, activation='lrelu')
        self.dropout = torch.nn.Dropout(p=dropout_rate)

        self.layer_names = []
        for idx in range(self.num_layers + 1):
            prev = max(idx - 1, 0)
            layer = SynthesisLayer(
                w_dim=self.w_dim, global_w_dim=512 * 2, is_torgb=(idx == self.num_layers),
                is_critically_sampled=(idx >= self.num_layers - self.num_critical), use_fp16=False,
                in_channels=int(channels[prev]), out_channels=int(channels[idx]),
                in_size=int(sizes[prev]), out_size=int(sizes[idx]),
                in_sampling_rate=int(sampling_rates[prev]), out_sampling_rate=int(sampling_rates[idx]),
                in_cutoff=cutoffs[prev], out_cutoff=cutoffs[idx],
                in_half_width=half_widths[prev], out_half_width=half_widths[idx], **layer_kwargs)
            name = f'L{idx}_{layer.out_size[0]}_{layer.out_channels}'
            setattr(self, name, layer)
            self.layer_names.append(name)

    def pool(self, x):
        return _AvgPool4Fn.apply(x)

    _u8_lut = None

    @classmethod
    def u8_lut(cls):
        """float32 value of every uint8 code under the reference's input transform: Normalize(0, 255) then
        clip(2 v - 1, -1, 1) in float64, cast to float32 (data/augment/transforms.py:604-616)."""
        if cls._u8_lut is None:
            v = np.arange(256, dtype=np.float64) / 255.0
            cls._u8_lut = np.ascontiguousarray(np.clip(2.0 * v - 1.0, -1, 1).astype(np.float32))
        return cls._u8_lut

    def pad_input(self, img):
        """NET:669 zero margin.  A uint8 image stack is normalised on the fly (the volume is uploaded as bytes:
        4x fewer host-to-device bytes than the float32 stack the reference uploads)."""
        N, C, H, W = img.shape
        m = self.margin_size
        y = torch.empty([N, C, H + 2 * m, W + 2 * m], dtype=torch.float32, device=img.device)
        lut = None
        if img.dtype == torch.uint8:
            lut = _lib.np_ptr(self.u8_lut())
            img = img.contiguous()
        else:
            img = img.float().contiguous()
        _lib.check(_lib.lib().afcm_pad_input(_lib.ptr(img), _lib.ptr(y), lut, N * C, H, W, m, _lib.stream_ptr(img.device)))
        return y

    def _grouped_affines(self, wcat):
        """styles of every synthesis layer from wcat [L, N, w_dim + global_w_dim] in ONE launch (afcm_fully_connected_grouped): the
        affine layers depend on the mapped styles only (NET:349-352), ToRGB's 1/sqrt(C k^2) style scale (NET:353-355) included.
        Returns None (the layers then run their own affine) when the layers do not share one input width."""
        import ctypes
        layers = [getattr(self, n) for n in self.layer_names]
        L_, N, in_f = wcat.shape
        if L_ > 16 or any(l.affine.in_features != in_f or l.affine.activation != 'linear' or l.affine.bias is None for l in layers):
            return None
        a0 = layers[0].affine
        if any(l.affine.weight_gain != a0.weight_gain or l.affine.bias_gain != a0.bias_gain for l in layers):
            return None
        outs = [torch.empty([N, l.affine.out_features], dtype=torch.float32, device=wcat.device) for l in layers]
        vp = ctypes.c_void_p * L_
        xs = vp(*[wcat[i].data_ptr() for i in range(L_)])
        wp = vp(*[l.affine.weight.data_ptr() for l in layers])
        bp = vp(*[l.affine.bias.data_ptr() for l in layers])
        yp = vp(*[o.data_ptr() for o in outs])
        of = (ctypes.c_int * L_)(*[l.affine.out_features for l in layers])
        gains = (ctypes.c_float * L_)(*[(1 / np.sqrt(l.in_channels * (l.conv_kernel ** 2))) if l.is_torgb else 1.0 for l in layers])
        _lib.check(_lib.lib().afcm_fully_connected_grouped(L_, xs, in_f, wp, bp, yp, of, gains, N, in_f, float(a0.weight_gain),
                                                           float(a0.bias_gain), _lib.stream_ptr(wcat.device)))
        return outs

    def forward(self, ws, img_in, **layer_kwargs):
        misc.assert_shape(ws, [None, self.num_ws, self.w_dim])
        _lib.require_cuda(ws, img_in)
        ws = ws.to(torch.float32).unbind(dim=1)
        x = self.pad_input(img_in)                                                    # NET:669
        E_features = {}
        # the reference calls the encoder layers without keyword arguments (NET:673-680): their magnitude_ema buffers are
        # never updated (they stay at 1.0 in reference checkpoints) -- only force_fp32 is forwarded here
        enc_kwargs = {k: v for k, v in layer_kwargs.items() if k in ('force_fp32',)}
        for idx in range(self.num_layers):                                            # NET:673-680
            rev_idx = self.num_layers - idx - 1
            rev_prev = self.num_layers - max(idx - 1, 0) - 1
            x = getattr(self, f'encoder_{idx}')(x, conv_ready=True, **enc_kwargs)     # every encoder output feeds a 3x3 convolution
            if (self.sizes[rev_idx] != self.sizes[rev_prev]) and self.sizes[rev_prev] != self.sizes[0]:
                E_features[self.sizes[rev_idx]] = x
        g = self.e_16x16(x)                                                           # NET:682-686
        g = self.pool(g)
        g = self.fc_in(g.flatten(1))
        img_global = self.dropout(g)
        res_idx = 1
        last = len(self.layer_names) - 1
        ws_layers = ws[1:]
        pre_cat = not torch.is_grad_enabled() and all(getattr(self, n).cond_mod for n in self.layer_names)
        styles_all = None
        if pre_cat:
            # [w_l, img_global] of every layer in ONE concatenation (the reference concatenates inside each layer, NET:349-352)
            L_ = len(self.layer_names)
            wcat = torch.cat([torch.stack(ws_layers[:L_], 0), img_global.unsqueeze(0).expand(L_, -1, -1)], dim=2)
            ws_layers = wcat.unbind(0)
            styles_all = self._grouped_affines(wcat)
        for idx, (name, w) in enumerate(zip(self.layer_names, ws_layers)):            # NET:691-698
            nxt = min(idx + 1, last)
            if (self.sizes[idx] != self.sizes[nxt]) and self.sizes[idx] != self.sizes[0]:
                include_skip = self.skip_connects[res_idx]
                res_idx += 1
            else:
                include_skip = False
            scale = self.output_scale if idx == last else 1.0                         # NET:699-700 folded
            # fast path: activations stay fp16 up to and including the input of the fused ToRGB kernel
            od = None
            feeds_conv = idx + 1 <= last and getattr(self, self.layer_names[idx + 1]).conv_kernel == 3
            x = getattr(self, name)(x, w, None if pre_cat else img_global, E_features, include_skip, out_scale=scale, out_dtype=od,
                                    conv_ready=feeds_conv, styles=None if styles_all is None else styles_all[idx], **layer_kwargs)
        misc.assert_shape(x, [None, self.img_channels_out, self.img_resolution, self.img_resolution])
        return x.to(torch.float32)


class Stylegan3Generator(torch.nn.Module):
    """NET:717-740."""

    def __init__(self, z_dim, c_dim, w_dim, img_resolution, img_channels_in, img_channels_out, mapping_kwargs={},
                 synthesis_kwargs={}):
        super().__init__()
        self.z_dim, self.c_dim, self.w_dim = z_dim, c_dim, w_dim
        self.img_resolution = img_resolution
        self.synthesis = SynthesisNetwork(w_dim=w_dim, img_resolution=img_resolution, img_channels_in=img_channels_in,
                                          img_channels_out=img_channels_out, **synthesis_kwargs)
        self.num_ws = self.synthesis.num_ws
        self.mapping = MappingNetwork(z_dim=z_dim, c_dim=c_dim, w_dim=w_dim, num_ws=self.num_ws, **mapping_kwargs)

    def forward(self, z, c, cond_img, ref_img=None, truncation_psi=1, truncation_cutoff=None, update_emas=False,
                **synthesis_kwargs):
        ws = self.mapping(z, c, img_in=ref_img, truncation_psi=truncation_psi, truncation_cutoff=truncation_cutoff,
                          update_emas=update_emas)
        return self.synthesis(ws, cond_img, update_emas=update_emas, **synthesis_kwargs)


def afcm_generator(seed=0, device='cuda', **overrides):
    """The generator every shipped stylegan3 config resolves to (models/stylegan3_model.py:37-65 with
    configs/adni/stylegan3/cmsr.yml:6-15), random-initialised like the reference."""
    cfg = dict(z_dim=512, c_dim=1, w_dim=512, img_resolution=256, img_channels_in=4, img_channels_out=1)
    syn = dict(channel_base=16384, channel_max=512, num_layers=14, num_critical=2, first_cutoff=2,
               first_stopband=2 ** 2.1, last_stopband_rel=2 ** 0.3, margin_size=10, output_scale=0.25,
               skip_resolution=128, conv_kernel=3, filter_size=6, lrelu_upsampling=2, use_radial_filters=False,
               conv_clamp=256, magnitude_ema_beta=0.5 ** (16 / (20 * 1e3)), cond_mod=True)
    mapping = dict(num_layers=8)
    for k, v in overrides.items():
        if k in cfg:
            cfg[k] = v
        elif k == 'mapping_layers':
            mapping['num_layers'] = v
        else:
            syn[k] = v
    torch.manual_seed(seed)
    G = Stylegan3Generator(mapping_kwargs=mapping, synthesis_kwargs=syn, **cfg).eval()
    return G.to(device) if device is not None else G
