"""CPU oracle for the AFCM generator hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module, and only as the checker or as the CPU baseline.  Nothing under afcm_b200/ imports it.

Two layers:
  * ctypes bindings to oracle/libafcm_oracle.so (plain C restatement, see afcm_oracle.c) -- used for
    the per-operator checks, including the sign tensor of the native filtered_lrelu.
  * a torch-CPU restatement of the whole generator forward (`generator_forward`) that follows
    models/networks/stylegan3/networks_stylegan3.py line by line (citations inline, paths relative to
    /root/reference; NET = that file, OPS = models/networks/stylegan3/torch_utils/ops).  It uses the
    same third-party arithmetic the reference's CPU `_ref` path uses (torch conv2d / matmul,
    scipy.signal.firwin), so it doubles as the "port" CPU baseline in bench.py.

Parity pinning: the reference has no tests / golden vectors.  This oracle is
pinned against the reference's own `_ref` implementation, imported and executed in the authoring
container by tests/golden/gen_golden.py; the resulting vectors are committed under tests/golden/ and
checked by tests/test_oracle_golden.py.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

ACT_IDX = dict(linear=1, relu=2, lrelu=3, tanh=4, sigmoid=5, elu=6, selu=7, softplus=8, swish=9)
ACT_DEF_ALPHA = dict(linear=0, relu=0, lrelu=0.2, tanh=0, sigmoid=0, elu=0, selu=0, softplus=0, swish=0)
ACT_DEF_GAIN = dict(linear=1, relu=np.sqrt(2), lrelu=np.sqrt(2), tanh=1, sigmoid=1, elu=1, selu=1, softplus=1,
                    swish=np.sqrt(2))


def build(force=False):
    """Compile oracle/libafcm_oracle.so with gcc (no CUDA involved)."""
    so = os.path.join(_HERE, 'libafcm_oracle.so')
    src = os.path.join(_HERE, 'afcm_oracle.c')
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(['make', '-C', _HERE, '-B', 'libafcm_oracle.so'], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
    return _LIB


def _fp(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def _f32(a):
    return None if a is None else np.ascontiguousarray(np.asarray(a, dtype=np.float32))


# ----------------------------------------------------------------------------------------------
# ctypes wrappers (numpy in, numpy out)

def bias_act(x, b=None, dim=1, act='linear', alpha=None, gain=None, clamp=None):
    x = _f32(x); b = _f32(b)
    alpha = float(ACT_DEF_ALPHA[act] if alpha is None else alpha)
    gain = float(ACT_DEF_GAIN[act] if gain is None else gain)
    clamp = float(-1 if clamp is None else clamp)
    y = np.empty_like(x)
    step = int(np.prod(x.shape[dim + 1:])) if b is not None else 1
    size = int(x.shape[dim]) if b is not None else 1
    rc = lib().orc_bias_act(_fp(x), _fp(b), None, None, _fp(y), ctypes.c_int64(x.size), ctypes.c_int64(step),
                            ctypes.c_int64(size), 0, ACT_IDX[act], ctypes.c_float(alpha), ctypes.c_float(gain),
                            ctypes.c_float(clamp))
    assert rc == 0
    return y


def bias_act_grad(dy, b, xref, yref, dim=1, act='linear', alpha=None, gain=None, clamp=None):
    dy = _f32(dy); b = _f32(b); xref = _f32(xref); yref = _f32(yref)
    alpha = float(ACT_DEF_ALPHA[act] if alpha is None else alpha)
    gain = float(ACT_DEF_GAIN[act] if gain is None else gain)
    clamp = float(-1 if clamp is None else clamp)
    dx = np.empty_like(dy)
    step = int(np.prod(dy.shape[dim + 1:])) if b is not None else 1
    size = int(dy.shape[dim]) if b is not None else 1
    rc = lib().orc_bias_act(_fp(dy), _fp(b), _fp(xref), _fp(yref), _fp(dx), ctypes.c_int64(dy.size),
                            ctypes.c_int64(step), ctypes.c_int64(size), 1, ACT_IDX[act], ctypes.c_float(alpha),
                            ctypes.c_float(gain), ctypes.c_float(clamp))
    assert rc == 0
    return dx


def _pad4(padding):
    if isinstance(padding, (int, np.integer)):
        padding = [padding, padding]
    padding = [int(p) for p in padding]
    if len(padding) == 2:
        padding = [padding[0], padding[0], padding[1], padding[1]]
    return padding


def _pair(v):
    return (int(v), int(v)) if isinstance(v, (int, np.integer)) else (int(v[0]), int(v[1]))


def upfirdn2d(x, f, up=1, down=1, padding=0, flip_filter=False, gain=1):
    x = _f32(x)
    N, C, H, W = x.shape
    upx, upy = _pair(up); dnx, dny = _pair(down)
    px0, px1, py0, py1 = _pad4(padding)
    f = np.ones([1, 1], np.float32) if f is None else _f32(f)
    if f.ndim == 1:
        fh, fw = 0, f.shape[0]; fhh = fw
    else:
        fh, fw = f.shape; fhh = fh
    OW = (W * upx + px0 + px1 - fw + dnx) // dnx
    OH = (H * upy + py0 + py1 - fhh + dny) // dny
    y = np.empty([N, C, OH, OW], np.float32)
    rc = lib().orc_upfirdn2d(_fp(x), ctypes.c_int64(N * C), H, W, _fp(y), _fp(f), fh, fw, upx, upy, dnx, dny,
                             px0, px1, py0, py1, int(bool(flip_filter)), ctypes.c_float(gain))
    assert rc == 0
    return y


def filtered_lrelu_sizes(H, W, up, down, fu_n, fd_n, padding):
    px0, px1, py0, py1 = _pad4(padding)
    v = [ctypes.c_int() for _ in range(6)]
    rc = lib().orc_filtered_lrelu_sizes(H, W, up, down, fu_n, fd_n, px0, px1, py0, py1, *[ctypes.byref(t) for t in v])
    assert rc == 0
    return dict(zip(['UH', 'UW', 'OH', 'OW', 'SH', 'SWB'], [t.value for t in v]))


def filtered_lrelu(x, fu=None, fd=None, b=None, up=1, down=1, padding=0, gain=np.sqrt(2), slope=0.2, clamp=None,
                   flip_filter=False, write_signs=False, si=None, sx=0, sy=0, return_preact=False):
    """Separable-filter filtered_lrelu.  Returns y, or (y, signs) / (y, signs, preact)."""
    x = _f32(x); b = _f32(b); fu = _f32(fu); fd = _f32(fd)
    N, C, H, W = x.shape
    fu_n = 1 if fu is None else fu.shape[0]
    fd_n = 1 if fd is None else fd.shape[0]
    assert fu is None or fu.ndim == 1
    assert fd is None or fd.ndim == 1
    px0, px1, py0, py1 = _pad4(padding)
    sz = filtered_lrelu_sizes(H, W, up, down, fu_n, fd_n, [px0, px1, py0, py1])
    y = np.empty([N, C, sz['OH'], sz['OW']], np.float32)
    so = np.zeros([N, C, sz['SH'], sz['SWB']], np.uint8) if write_signs else None
    pre = np.empty([N, C, sz['UH'], sz['UW']], np.float32) if return_preact else None
    si_h = si_wb = 0
    if si is not None:
        si = np.ascontiguousarray(si, dtype=np.uint8); si_h, si_wb = si.shape[2], si.shape[3]
    clamp_v = float('inf') if clamp is None else float(clamp)
    rc = lib().orc_filtered_lrelu(_fp(x), ctypes.c_int64(N), ctypes.c_int64(C), H, W, _fp(b), _fp(fu), fu_n, _fp(fd),
                                  fd_n, up, down, px0, px1, py0, py1, ctypes.c_float(gain), ctypes.c_float(slope),
                                  ctypes.c_float(clamp_v), int(bool(flip_filter)), _fp(y), _fp(so), _fp(si), si_h,
                                  si_wb, int(sx), int(sy), _fp(pre))
    assert rc == 0
    out = [y]
    if write_signs:
        out.append(so)
    if return_preact:
        out.append(pre)
    return out[0] if len(out) == 1 else tuple(out)


def conv2d(x, w, padding=0):
    x = _f32(x); w = _f32(w)
    N, Ci, H, W = x.shape; Co, Ci2, kh, kw = w.shape
    assert Ci == Ci2
    y = np.empty([N, Co, H + 2 * padding - kh + 1, W + 2 * padding - kw + 1], np.float32)
    rc = lib().orc_conv2d(_fp(x), ctypes.c_int64(N), Ci, H, W, _fp(w), Co, kh, kw, int(padding), _fp(y))
    assert rc == 0
    return y


def modulated_conv2d(x, w, s, demodulate=True, padding=0, input_gain=None):
    x = _f32(x); w = _f32(w); s = _f32(s)
    N, Ci, H, W = x.shape; Co, _, kh, kw = w.shape
    g = None if input_gain is None else _f32(np.asarray(input_gain, np.float32).reshape(-1))
    y = np.empty([N, Co, H + 2 * padding - kh + 1, W + 2 * padding - kw + 1], np.float32)
    rc = lib().orc_modulated_conv2d(_fp(x), ctypes.c_int64(N), Ci, H, W, _fp(w), Co, kh, kw, _fp(s),
                                    int(bool(demodulate)), int(padding), _fp(g),
                                    ctypes.c_int64(0 if g is None else g.size), _fp(y))
    assert rc == 0
    return y


def fully_connected(x, w, b=None, weight_gain=1.0, bias_gain=1.0, act='linear'):
    x = _f32(x); w = _f32(w); b = _f32(b)
    y = np.empty([x.shape[0], w.shape[0]], np.float32)
    rc = lib().orc_fully_connected(_fp(x), ctypes.c_int64(x.shape[0]), x.shape[1], _fp(w), w.shape[0], _fp(b),
                                   ctypes.c_float(weight_gain), ctypes.c_float(bias_gain), ACT_IDX[act],
                                   ctypes.c_float(ACT_DEF_ALPHA[act]), ctypes.c_float(ACT_DEF_GAIN[act]), _fp(y))
    assert rc == 0
    return y


def num_threads():
    return int(lib().orc_num_threads())


# ----------------------------------------------------------------------------------------------
# Whole-generator restatement on torch CPU tensors.

DEFAULT_CFG = dict(  # models/stylegan3_model.py:37-65 + configs/adni/stylegan3/cmsr.yml:6-15
    z_dim=512, c_dim=1, w_dim=512, img_resolution=256, img_channels_in=4, img_channels_out=1,
    mapping_layers=8, channel_base=16384, channel_max=512, num_layers=14, num_critical=2, first_cutoff=2,
    first_stopband=2 ** 2.1, last_stopband_rel=2 ** 0.3, margin_size=10, output_scale=0.25, skip_resolution=128,
    conv_kernel=3, filter_size=6, lrelu_upsampling=2, conv_clamp=256)


def _lowpass(numtaps, cutoff, width, fs):
    """NET:381-392 -- Kaiser low-pass via scipy.signal.firwin; None for the 1-tap identity."""
    import scipy.signal
    if numtaps == 1:
        return None
    return np.asarray(scipy.signal.firwin(numtaps=numtaps, cutoff=cutoff, width=width, fs=fs), dtype=np.float32)


def _one_layer(in_ch, out_ch, in_size, out_size, in_sr, out_sr, in_cut, out_cut, in_hw, out_hw, cfg, torgb, k):
    """Filter design and padding of one Encoder/Synthesis layer (NET:294-334 and NET:453-489)."""
    tmp_sr = max(in_sr, out_sr) * (1 if torgb else cfg['lrelu_upsampling'])
    up = int(np.rint(tmp_sr / in_sr)); down = int(np.rint(tmp_sr / out_sr))
    up_taps = cfg['filter_size'] * up if (up > 1 and not torgb) else 1
    down_taps = cfg['filter_size'] * down if (down > 1 and not torgb) else 1
    total = (out_size - 1) * down + 1 - (in_size + k - 1) * up + up_taps + down_taps - 2
    lo = (total + up) // 2
    return dict(in_channels=in_ch, out_channels=out_ch, in_size=in_size, out_size=out_size, up=up, down=down,
                up_taps=up_taps, down_taps=down_taps, padding=[lo, total - lo, lo, total - lo], is_torgb=torgb,
                conv_kernel=k, up_filter=_lowpass(up_taps, in_cut, in_hw * 2, tmp_sr),
                down_filter=_lowpass(down_taps, out_cut, out_hw * 2, tmp_sr))


def layer_specs(cfg=None):
    """Geometric schedule of NET:595-664.  Returns (encoder_specs, synthesis_specs, sizes, skip_connects)."""
    cfg = dict(DEFAULT_CFG, **(cfg or {}))
    L = cfg['num_layers']; res = cfg['img_resolution']
    last_cutoff = res / 2
    last_stopband = last_cutoff * cfg['last_stopband_rel']
    e = np.minimum(np.arange(L + 1) / (L - cfg['num_critical']), 1)
    cut = cfg['first_cutoff'] * (last_cutoff / cfg['first_cutoff']) ** e
    stop = cfg['first_stopband'] * (last_stopband / cfg['first_stopband']) ** e
    sr = np.exp2(np.ceil(np.log2(np.minimum(stop * 2, res))))
    hw = np.maximum(stop, sr / 2) - cut
    sizes = sr + cfg['margin_size'] * 2
    enc_sizes = sizes.copy()
    sizes[-2:] = res
    ch = np.rint(np.minimum((cfg['channel_base'] / 2) / cut, cfg['channel_max']))
    ch[-1] = cfg['img_channels_out']
    k = cfg['conv_kernel']
    enc = []
    for i in range(L):
        r = L - i - 1; pv = max(i - 1, 0); rp = L - pv - 1
        cin = cfg['img_channels_in'] if i == 0 else int(ch[rp])
        enc.append(_one_layer(cin, int(ch[r]), int(enc_sizes[rp]), int(enc_sizes[r]), int(sr[rp]), int(sr[r]),
                              cut[rp], cut[r], hw[rp], hw[r], cfg, False, k))
    syn = []
    for i in range(L + 1):
        pv = max(i - 1, 0); torgb = (i == L)
        sp = _one_layer(int(ch[pv]), int(ch[i]), int(sizes[pv]), int(sizes[i]), int(sr[pv]), int(sr[i]),
                        cut[pv], cut[i], hw[pv], hw[i], cfg, torgb, 1 if torgb else k)
        sp['name'] = f"L{i}_{int(sizes[i])}_{int(ch[i])}"
        syn.append(sp)
    sk = cfg['skip_resolution']
    log2res = int(np.log2(res))
    if sk >= 4:
        fs = int(np.log2(sk))
        skips = [True] * (fs - 1) + [False] * (log2res - fs)
    else:
        skips = [False] * log2res
    return enc, syn, sizes, skips, int(ch[0])


def _t_upfirdn(x, f, up, down, pad, gain):
    """OPS/upfirdn2d.py:167-211 for a separable 1-D filter (or None), same factor on both axes."""
    import torch
    import torch.nn.functional as F
    N, C, H, W = x.shape
    if up > 1:
        z = x.new_zeros(N, C, H * up, W * up)
        z[:, :, ::up, ::up] = x
        x = z
    px0, px1, py0, py1 = pad
    x = F.pad(x, [max(px0, 0), max(px1, 0), max(py0, 0), max(py1, 0)])
    x = x[:, :, max(-py0, 0): x.shape[2] - max(-py1, 0), max(-px0, 0): x.shape[3] - max(-px1, 0)]
    if f is None:
        x = x * gain if gain != 1 else x
    else:
        k = (torch.as_tensor(f) * (gain ** 0.5)).flip(0)          # true convolution
        kw = k.reshape(1, 1, 1, -1).repeat(C, 1, 1, 1)
        kh = k.reshape(1, 1, -1, 1).repeat(C, 1, 1, 1)
        x = F.conv2d(x, kw, groups=C)
        x = F.conv2d(x, kh, groups=C)
    return x[:, :, ::down, ::down]


def t_filtered_lrelu(x, fu, fd, b, up, down, padding, gain, slope, clamp):
    """OPS/filtered_lrelu.py:121-153 on torch CPU tensors."""
    import torch.nn.functional as F
    x = x + b.reshape(1, -1, 1, 1)                                  # :145
    x = _t_upfirdn(x, fu, up, 1, padding, up ** 2)                  # :146
    x = F.leaky_relu(x, slope) * gain                               # :147 (bias_act 'lrelu': act, gain, clamp)
    if clamp is not None:
        x = x.clamp(-clamp, clamp)
    return _t_upfirdn(x, fd, 1, down, [0, 0, 0, 0], 1)              # :148


def t_modulated_conv2d(x, w, s, demodulate, padding, input_gain):
    """NET:25-64."""
    import torch.nn.functional as F
    B = x.shape[0]; O, I, kh, kw = w.shape
    if demodulate:
        w = w * w.square().mean([1, 2, 3], keepdim=True).rsqrt()
        s = s * s.square().mean().rsqrt()
    wm = w[None] * s[:, None, :, None, None]
    if demodulate:
        wm = wm * (wm.square().sum([2, 3, 4], keepdim=True) + 1e-8).rsqrt()
    if input_gain is not None:
        wm = wm * input_gain.expand(B, I)[:, None, :, None, None]
    y = F.conv2d(x.reshape(1, B * I, *x.shape[2:]), wm.reshape(B * O, I, kh, kw), padding=padding, groups=B)
    return y.reshape(B, O, *y.shape[2:])


def t_fc(x, weight, bias, lr_mul=1.0, act='linear'):
    """NET:89-101."""
    import torch.nn.functional as F
    w = weight * (lr_mul / np.sqrt(weight.shape[1]))
    y = x.matmul(w.t())
    if bias is not None:
        y = y + bias * lr_mul
    if act == 'lrelu':
        y = F.leaky_relu(y, 0.2) * np.sqrt(2)
    else:
        assert act == 'linear'
    return y


def mapping_forward(P, z, c, cfg=None):
    """NET:135-161 (truncation_psi == 1, update_emas False)."""
    import torch
    cfg = dict(DEFAULT_CFG, **(cfg or {}))
    x = z.to(torch.float32)
    x = x * (x.square().mean(1, keepdim=True) + 1e-8).rsqrt()
    if cfg['c_dim'] > 0:
        y = t_fc(c.to(torch.float32), P['mapping.embed.weight'], P['mapping.embed.bias'])
        y = y * (y.square().mean(1, keepdim=True) + 1e-8).rsqrt()
        x = torch.cat([x, y], 1)
    for i in range(cfg['mapping_layers']):
        x = t_fc(x, P[f'mapping.fc{i}.weight'], P[f'mapping.fc{i}.bias'], lr_mul=0.01, act='lrelu')
    return x.unsqueeze(1).repeat(1, cfg['num_layers'] + 2, 1)


def synthesis_forward(P, ws, img, cfg=None, taps=None):
    """NET:666-705.  `taps` (optional dict) receives intermediate activations for layer-wise checks."""
    import torch
    import torch.nn.functional as F
    cfg = dict(DEFAULT_CFG, **(cfg or {}))
    enc, syn, sizes, skips, _ = layer_specs(cfg)
    L = cfg['num_layers']; m = cfg['margin_size']; clamp = cfg['conv_clamp']
    x = F.pad(img.to(torch.float32), [m] * 4)                                            # :669
    feats = {}
    for i, sp in enumerate(enc):                                                          # :673-680
        pre = f'synthesis.encoder_{i}'
        w = P[pre + '.weight'] * (1 / np.sqrt(sp['in_channels'] * sp['conv_kernel'] ** 2))  # :503
        x = F.conv2d(x, w, padding=sp['conv_kernel'] - 1)                                 # :505
        x = t_filtered_lrelu(x, sp['up_filter'], sp['down_filter'], P[pre + '.bias'], sp['up'], sp['down'],
                             sp['padding'], np.sqrt(2), 0.2, clamp)                       # :510-511
        r = L - i - 1; rp = L - max(i - 1, 0) - 1
        if sizes[r] != sizes[rp] and sizes[rp] != sizes[0]:
            feats[int(sizes[r])] = x
        if taps is not None:
            taps[f'enc{i}'] = x
    # bottleneck -> global code                                                           # :682-686
    w = P['synthesis.e_16x16.weight'] * (1 / np.sqrt(P['synthesis.e_16x16.weight'][0].numel()))
    g = F.conv2d(x, w, padding=1) + P['synthesis.e_16x16.bias'].reshape(1, -1, 1, 1)
    g = F.leaky_relu(g, 0.2) * np.sqrt(2)
    g = F.adaptive_avg_pool2d(g, (4, 4)).flatten(1)
    g = t_fc(g, P['synthesis.fc_in.weight'], P['synthesis.fc_in.bias'], act='lrelu')      # dropout = identity (eval)
    if taps is not None:
        taps['global'] = g
    res_idx = 1
    for i, sp in enumerate(syn):                                                          # :691-698
        pre = 'synthesis.' + sp['name']
        nxt = min(i + 1, len(syn) - 1)
        if sizes[i] != sizes[nxt] and sizes[i] != sizes[0]:
            skip = skips[res_idx]; res_idx += 1
        else:
            skip = False
        gain_in = P[pre + '.magnitude_ema'].rsqrt()                                        # :346
        st = t_fc(torch.cat([ws[:, i + 1], g], 1), P[pre + '.affine.weight'], P[pre + '.affine.bias'])  # :349-352
        if sp['is_torgb']:
            st = st * (1 / np.sqrt(sp['in_channels'] * sp['conv_kernel'] ** 2))           # :353-355
        x = t_modulated_conv2d(x, P[pre + '.weight'], st, not sp['is_torgb'], sp['conv_kernel'] - 1, gain_in)
        x = t_filtered_lrelu(x, sp['up_filter'], sp['down_filter'], P[pre + '.bias'], sp['up'], sp['down'],
                             sp['padding'], 1 if sp['is_torgb'] else np.sqrt(2), 1 if sp['is_torgb'] else 0.2, clamp)
        if skip:
            x = x + feats[sp['out_size']]                                                  # :360-363,376-377
        if taps is not None:
            taps[sp['name']] = x
    if cfg['output_scale'] != 1:
        x = x * cfg['output_scale']                                                        # :699-700
    return x


def generator_forward(P, z, c, cond_img, cfg=None, taps=None, grad=False):
    """Stylegan3Generator.forward, NET:737-740 (eval mode, noise_mode='const', truncation_psi=1).  grad=True keeps the
    autograd graph: the restatement is built from differentiable tensor operations only, so torch's autograd over it is
    the CPU checker of the training-step path (pinned by tests/golden/tiny_gen_grads.npz)."""
    import torch
    with torch.set_grad_enabled(bool(grad)):
        ws = mapping_forward(P, z, c, cfg)
        return synthesis_forward(P, ws, cond_img, cfg, taps)


def init_params(cfg=None, seed=0):
    """Random-init parameters drawn in the SAME ORDER as the reference constructors consume the torch
    RNG (NET:625-664 encoder layers, e_16x16, fc_in, synthesis layers [affine, weight]; NET:128-132
    mapping embed, fc0..), so torch.manual_seed(seed) reproduces the reference's random-init weights."""
    import torch
    cfg = dict(DEFAULT_CFG, **(cfg or {}))
    enc, syn, _, _, ch0 = layer_specs(cfg)
    torch.manual_seed(seed)
    P = {}
    k = cfg['conv_kernel']
    for i, sp in enumerate(enc):
        pre = f'synthesis.encoder_{i}'
        P[pre + '.weight'] = torch.randn([sp['out_channels'], sp['in_channels'], k, k])
        P[pre + '.bias'] = torch.zeros([sp['out_channels']])
        P[pre + '.magnitude_ema'] = torch.ones([])
    P['synthesis.e_16x16.weight'] = torch.randn([ch0, ch0, 3, 3])
    P['synthesis.e_16x16.bias'] = torch.zeros([ch0])
    P['synthesis.fc_in.weight'] = torch.randn([1024, ch0 * 16])
    P['synthesis.fc_in.bias'] = torch.zeros([1024])
    for sp in syn:
        pre = 'synthesis.' + sp['name']
        P[pre + '.affine.weight'] = torch.randn([sp['in_channels'], cfg['w_dim'] + 1024])
        P[pre + '.affine.bias'] = torch.ones([sp['in_channels']])
        P[pre + '.weight'] = torch.randn([sp['out_channels'], sp['in_channels'], sp['conv_kernel'], sp['conv_kernel']])
        P[pre + '.bias'] = torch.zeros([sp['out_channels']])
        P[pre + '.magnitude_ema'] = torch.ones([])
    if cfg['c_dim'] > 0:
        P['mapping.embed.weight'] = torch.randn([cfg['w_dim'], cfg['c_dim']])
        P['mapping.embed.bias'] = torch.zeros([cfg['w_dim']])
    feats = [cfg['z_dim'] + (cfg['w_dim'] if cfg['c_dim'] > 0 else 0)] + [cfg['w_dim']] * cfg['mapping_layers']
    for i in range(cfg['mapping_layers']):
        P[f'mapping.fc{i}.weight'] = torch.randn([feats[i + 1], feats[i]]) * (1 / 0.01)
        P[f'mapping.fc{i}.bias'] = torch.zeros([feats[i + 1]])
    return P
