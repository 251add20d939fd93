"""Generates tests/golden/*.npz by importing and running the REFERENCE's own Python `_ref` path
(CPU, float32) in the authoring container.  The reference ships no tests or golden vectors, so these
files are what pins both the oracle (oracle/) and the CUDA path.  /root/reference does not exist on
the GPU box -- only the committed .npz files travel.

    python tests/golden/gen_golden.py            # writes ops.npz, tiny_gen.npz, tiny_gen_grads.npz, full_gen.npz

Seeds are fixed; re-running reproduces the files bit-for-bit with the same torch build.
"""
import os
import sys
import warnings

sys.dont_write_bytecode = True
sys.path.insert(0, '/root/reference')
warnings.filterwarnings('ignore')

import numpy as np  # noqa: E402
import torch  # noqa: E402

from models.networks.stylegan3.networks_stylegan3 import (  # noqa: E402
    Stylegan3Generator, modulated_conv2d, FullyConnectedLayer, MappingNetwork, SynthesisInput)
from models.networks.stylegan3.torch_utils.ops import filtered_lrelu, upfirdn2d, bias_act  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(8)


def firwin_like(taps, cutoff, width, fs):
    import scipy.signal
    return torch.as_tensor(scipy.signal.firwin(numtaps=taps, cutoff=cutoff, width=width, fs=fs), dtype=torch.float32)


def gen_ops():
    out = {}
    g = torch.Generator().manual_seed(1234)

    def rn(*shape, scale=1.0):
        return torch.randn(*shape, generator=g) * scale

    # ---- filtered_lrelu, every geometry class the AFCM generator uses ----
    f12 = firwin_like(12, 128.0, 59.17, 512)      # enc0-like
    f12b = firwin_like(12, 5.657, 20.69, 64)
    f24 = firwin_like(24, 4.0, 8.0, 64)
    cases = [
        # name,   H,  W,  C, up, dn, fu,   fd,   padding,            gain,        slope, clamp, scale
        ('u2d2',  38, 38, 3, 2, 2, f12,  f12b, [9, 8, 9, 8],        np.sqrt(2),  0.2,   256,   1.0),
        ('u2d4',  54, 54, 2, 2, 4, f12b, f24,  [34, 33, 34, 33],    np.sqrt(2),  0.2,   256,   1.0),
        ('u4d2',  38, 38, 2, 4, 2, f24,  f12b, [-6, -9, -6, -9],    np.sqrt(2),  0.2,   256,   1.0),
        ('crop',  54, 54, 2, 2, 2, f12,  f12b, [-11, -12, -11, -12], np.sqrt(2), 0.2,   256,   1.0),
        ('torgb', 32, 32, 3, 1, 1, None, None, [0, 0, 0, 0],        1.0,         1.0,   256,   200.0),
        ('clamp', 38, 38, 2, 2, 2, f12,  f12b, [9, 8, 9, 8],        np.sqrt(2),  0.2,   1.5,   2.0),
        ('rect',  30, 46, 2, 2, 2, f12,  f12b, [9, 8, 7, 10],       1.3,         0.1,   None,  1.0),
        ('u2d1',  20, 20, 2, 2, 1, f12,  None, [5, 6, 5, 6],        np.sqrt(2),  0.2,   256,   1.0),
        ('u1d2',  40, 40, 2, 1, 2, None, f12b, [5, 6, 5, 6],        np.sqrt(2),  0.2,   256,   1.0),
    ]
    names = []
    for (name, H, W, C, up, dn, fu, fd, pad, gain, slope, clamp, scale) in cases:
        x = rn(2, C, H, W, scale=scale)
        x[0, 0, 3, 4:9] = 0.0                          # exact zeros exercise the sign bit convention
        b = rn(C, scale=0.5)
        y = filtered_lrelu.filtered_lrelu(x, fu=fu, fd=fd, b=b, up=up, down=dn, padding=pad, gain=gain, slope=slope,
                                          clamp=clamp, impl='ref')
        # flip_filter=True variant (what the backward pass of the native op uses)
        yf = filtered_lrelu.filtered_lrelu(x, fu=fu, fd=fd, b=b, up=up, down=dn, padding=pad, gain=gain, slope=slope,
                                           clamp=clamp, flip_filter=True, impl='ref')
        k = 'flrelu.' + name
        names.append(name)
        out[k + '.x'] = x.numpy(); out[k + '.b'] = b.numpy(); out[k + '.y'] = y.numpy(); out[k + '.yflip'] = yf.numpy()
        out[k + '.fu'] = fu.numpy() if fu is not None else np.zeros([0], np.float32)
        out[k + '.fd'] = fd.numpy() if fd is not None else np.zeros([0], np.float32)
        out[k + '.cfg'] = np.asarray([up, dn] + pad + [gain, slope, -1 if clamp is None else clamp], np.float64)
        # gradient wrt x and b of sum(y * r) through the reference's autograd (pins the backward semantics)
        xg = x.clone().requires_grad_(True); bg = b.clone().requires_grad_(True)
        r = rn(*y.shape)
        yy = filtered_lrelu.filtered_lrelu(xg, fu=fu, fd=fd, b=bg, up=up, down=dn, padding=pad, gain=gain, slope=slope,
                                           clamp=clamp, impl='ref')
        (yy * r).sum().backward()
        out[k + '.r'] = r.numpy(); out[k + '.dx'] = xg.grad.numpy(); out[k + '.db'] = bg.grad.numpy()
    out['flrelu.names'] = np.asarray(names)

    # ---- upfirdn2d ----
    f2d = torch.outer(torch.tensor([1., 3., 3., 1.]), torch.tensor([1., 3., 3., 1.])); f2d = f2d / f2d.sum()
    f61 = torch.as_tensor(np.exp(-0.5 * (np.arange(-30, 31) / 10.0) ** 2), dtype=torch.float32); f61 = f61 / f61.sum()
    ucases = [
        ('sep_up2',   f12,  2, 1, [9, 8, 9, 8], False, 4.0),
        ('sep_dn2',   f12b, 1, 2, [0, 0, 0, 0], False, 1.0),
        ('full_1331', f2d,  1, 2, [1, 1, 1, 1], False, 1.0),
        ('full_up2',  f2d,  2, 1, [2, 1, 2, 1], True,  4.0),
        ('blur61',    f61,  1, 1, [30, 30, 30, 30], False, 1.0),     # stylegan3_model.py:28 blur (filter2d)
        ('none',      None, 1, 1, [1, -2, 0, 3], False, 2.0),
        ('asym',      torch.tensor([[1., 2., 3.], [4., 5., 6.]]) / 21, [2, 1], [1, 2], [2, 1, 0, 3], False, 1.5),
    ]
    unames = []
    for (name, f, up, dn, pad, flip, gain) in ucases:
        x = rn(2, 3, 24, 28)
        y = upfirdn2d.upfirdn2d(x, f, up=up, down=dn, padding=pad, flip_filter=flip, gain=gain, impl='ref')
        k = 'upfirdn.' + name
        unames.append(name)
        upx, upy = (up, up) if isinstance(up, int) else up
        dnx, dny = (dn, dn) if isinstance(dn, int) else dn
        out[k + '.x'] = x.numpy(); out[k + '.y'] = y.numpy()
        out[k + '.f'] = f.numpy() if f is not None else np.zeros([0], np.float32)
        out[k + '.cfg'] = np.asarray([upx, upy, dnx, dny] + pad + [int(flip), gain], np.float64)
    out['upfirdn.names'] = np.asarray(unames)

    # ---- bias_act: all 9 activations, with and without clamp ----
    x = rn(3, 5, 6, 7, scale=2.0); b = rn(5)
    out['bias_act.x'] = x.numpy(); out['bias_act.b'] = b.numpy()
    for act in bias_act.activation_funcs:
        for clamp in (None, 0.7):
            xg = x.clone().requires_grad_(True)
            y = bias_act.bias_act(xg, b, act=act, clamp=clamp, impl='ref')
            dy = torch.ones_like(y) * 0.5
            y.backward(dy)
            tag = f'bias_act.{act}.{"c" if clamp else "n"}'
            out[tag + '.y'] = y.detach().numpy(); out[tag + '.dx'] = xg.grad.numpy()
    x2 = rn(4, 9); b2 = rn(9)
    out['bias_act.x2'] = x2.numpy(); out['bias_act.b2'] = b2.numpy()
    out['bias_act.y2'] = bias_act.bias_act(x2, b2, act='lrelu', impl='ref').numpy()

    # ---- modulated_conv2d ----
    for name, (N, I, O, H, k, demod) in dict(demod3=(3, 10, 7, 13, 3, True), torgb1=(2, 16, 1, 12, 1, False),
                                             demod3b=(2, 24, 20, 38, 3, True)).items():
        x = rn(N, I, H, H); w = rn(O, I, k, k); s = rn(N, I) + 1.0; ig = torch.tensor(0.8)
        y = modulated_conv2d(x=x, w=w, s=s, demodulate=demod, padding=k - 1, input_gain=ig)
        t = 'modconv.' + name
        out[t + '.x'] = x.numpy(); out[t + '.w'] = w.numpy(); out[t + '.s'] = s.numpy(); out[t + '.y'] = y.numpy()
        out[t + '.cfg'] = np.asarray([int(demod), k - 1, 0.8])
        # gradients of sum(y * r) through the reference's autograd (grouped-conv backward of NET:60-63)
        xg = x.clone().requires_grad_(True); wg = w.clone().requires_grad_(True); sg = s.clone().requires_grad_(True)
        r = rn(*y.shape)
        (modulated_conv2d(x=xg, w=wg, s=sg, demodulate=demod, padding=k - 1, input_gain=ig) * r).sum().backward()
        out[t + '.r'] = r.numpy(); out[t + '.dx'] = xg.grad.numpy(); out[t + '.dw'] = wg.grad.numpy(); out[t + '.ds'] = sg.grad.numpy()

    # ---- FullyConnectedLayer / MappingNetwork / SynthesisInput ----
    torch.manual_seed(7)
    fc = FullyConnectedLayer(37, 19, activation='lrelu', lr_multiplier=0.01)
    with torch.no_grad():
        fc.bias.copy_(torch.randn(19) * 10)
    x = rn(5, 37)
    out['fc.x'] = x.numpy(); out['fc.w'] = fc.weight.detach().numpy(); out['fc.b'] = fc.bias.detach().numpy()
    out['fc.y'] = fc(x).detach().numpy()
    fc2 = FullyConnectedLayer(37, 19, bias_init=1)
    out['fc2.w'] = fc2.weight.detach().numpy(); out['fc2.b'] = fc2.bias.detach().numpy()
    out['fc2.y'] = fc2(x).detach().numpy()
    mp = MappingNetwork(z_dim=48, c_dim=1, w_dim=40, num_ws=6, num_layers=3)
    z = rn(4, 48); c = torch.rand(4, 1, generator=g)
    out['map.z'] = z.numpy(); out['map.c'] = c.numpy(); out['map.ws'] = mp(z, c).detach().numpy()
    for kk, v in mp.state_dict().items():
        out['map.P.' + kk] = v.numpy()
    si = SynthesisInput(w_dim=40, channels=12, size=20, sampling_rate=16, bandwidth=2)
    with torch.no_grad():
        si.affine.weight.copy_(torch.randn_like(si.affine.weight) * 0.3)
    w = rn(3, 40)
    out['synin.w'] = w.numpy(); out['synin.y'] = si(w).detach().numpy()
    for kk, v in si.state_dict().items():
        out['synin.P.' + kk] = v.numpy()
    np.savez_compressed(os.path.join(HERE, 'ops.npz'), **out)
    print('ops.npz', len(out), 'arrays')


TINY = dict(z_dim=64, c_dim=1, w_dim=64, img_resolution=32, img_channels_in=4, img_channels_out=1,
            mapping_layers=3, channel_base=512, channel_max=48, num_layers=6, num_critical=2, first_cutoff=2,
            first_stopband=2 ** 2.1, last_stopband_rel=2 ** 0.3, margin_size=10, output_scale=0.25,
            skip_resolution=16, conv_kernel=3, filter_size=6, lrelu_upsampling=2, conv_clamp=256)
FULL = dict(TINY, z_dim=512, w_dim=512, img_resolution=256, mapping_layers=8, channel_base=16384, channel_max=512,
            num_layers=14, skip_resolution=128)


def build_ref(cfg, seed):
    torch.manual_seed(seed)
    return Stylegan3Generator(
        z_dim=cfg['z_dim'], c_dim=cfg['c_dim'], w_dim=cfg['w_dim'], img_resolution=cfg['img_resolution'],
        img_channels_in=cfg['img_channels_in'], img_channels_out=cfg['img_channels_out'],
        mapping_kwargs=dict(num_layers=cfg['mapping_layers']),
        synthesis_kwargs=dict(channel_base=cfg['channel_base'], channel_max=cfg['channel_max'],
                              num_layers=cfg['num_layers'], num_critical=cfg['num_critical'],
                              first_cutoff=cfg['first_cutoff'], first_stopband=cfg['first_stopband'],
                              last_stopband_rel=cfg['last_stopband_rel'], margin_size=cfg['margin_size'],
                              output_scale=cfg['output_scale'], skip_resolution=cfg['skip_resolution'],
                              conv_kernel=cfg['conv_kernel'], filter_size=cfg['filter_size'],
                              lrelu_upsampling=cfg['lrelu_upsampling'], use_radial_filters=False,
                              conv_clamp=cfg['conv_clamp'], magnitude_ema_beta=0.5 ** (16 / (20 * 1e3)),
                              cond_mod=True)).eval()


def hook_taps(G, taps):
    hs = []
    S = G.synthesis
    for i in range(S.num_layers):
        hs.append(getattr(S, f'encoder_{i}').register_forward_hook(
            lambda m, a, o, i=i: taps.__setitem__(f'enc{i}', o.detach().clone())))
    for n in S.layer_names:
        hs.append(getattr(S, n).register_forward_hook(lambda m, a, o, n=n: taps.__setitem__(n, o.detach().clone())))
    hs.append(S.fc_in.register_forward_hook(lambda m, a, o: taps.__setitem__('global', o.detach().clone())))
    return hs


def synth_inputs(cfg, B, seed):
    """uint8-quantised slices mapped to [-1,1] (data/augment/transforms.py:604-616), c in [0,1)."""
    g = torch.Generator().manual_seed(seed)
    res = cfg['img_resolution']
    u8 = torch.randint(0, 256, (B, cfg['img_channels_in'], res, res), generator=g, dtype=torch.uint8)
    x = (u8.float() * (2.0 / 255.0) - 1.0).clamp(-1, 1)
    z = torch.randn(B, cfg['z_dim'], generator=g)
    c = torch.randint(0, 5, (B, 1), generator=g).float() / 5.0
    return z, c, x, u8


def gen_tiny():
    G = build_ref(TINY, seed=0)
    g = torch.Generator().manual_seed(99)
    with torch.no_grad():                      # non-trivial biases / input gains so the test has teeth
        for n, p in G.named_parameters():
            if n.endswith('.bias') and 'affine' not in n:
                p.copy_(torch.randn(p.shape, generator=g) * 0.1)
        for n, buf in G.named_buffers():
            if n.endswith('magnitude_ema'):
                buf.copy_(torch.rand([], generator=g) * 1.5 + 0.5)
    z, c, x, u8 = synth_inputs(TINY, 3, seed=5)
    taps = {}
    hook_taps(G, taps)
    with torch.no_grad():
        y = G(z, c, x, noise_mode='const')
    out = {'z': z.numpy(), 'c': c.numpy(), 'x': x.numpy(), 'y': y.numpy()}
    for k, v in taps.items():                       # first sample / first 8 channels of every activation + stats
        out['tap.' + k] = (v[:1, :8] if v.ndim == 4 else v).numpy().copy()
        out['stat.' + k] = np.asarray([v.double().mean().item(), v.double().std().item(), v.abs().max().item()])
    for k, v in G.state_dict().items():
        if not (k.endswith('up_filter') or k.endswith('down_filter') or k.endswith('resample_filter')):
            out['P.' + k] = v.numpy()
        else:
            out['F.' + k] = v.numpy()
    np.savez_compressed(os.path.join(HERE, 'tiny_gen.npz'), **out)
    print('tiny_gen.npz', y.shape, float(y.abs().max()))


def gen_tiny_grads():
    """Gradients of the L1 training loss mean|G(z,c,x) - target| with respect to every generator parameter, from the
    reference's own autograd on its CPU `_ref` operators (pins the training-step path, BASELINE config 5)."""
    G = build_ref(TINY, seed=0)
    g = torch.Generator().manual_seed(99)
    with torch.no_grad():                      # same perturbation as gen_tiny
        for n, p in G.named_parameters():
            if n.endswith('.bias') and 'affine' not in n:
                p.copy_(torch.randn(p.shape, generator=g) * 0.1)
        for n, buf in G.named_buffers():
            if n.endswith('magnitude_ema'):
                buf.copy_(torch.rand([], generator=g) * 1.5 + 0.5)
    z, c, x, u8 = synth_inputs(TINY, 3, seed=5)
    tg = torch.Generator().manual_seed(17)
    target = torch.rand(3, 1, TINY['img_resolution'], TINY['img_resolution'], generator=tg) * 2 - 1
    for p in G.parameters():
        p.requires_grad_(True)
    y = G(z, c, x, noise_mode='const')
    loss = (y - target).abs().mean()
    loss.backward()
    out = {'target': target.numpy(), 'loss': np.asarray(loss.item())}
    for n, p in G.named_parameters():
        out['G.' + n] = (p.grad if p.grad is not None else torch.zeros_like(p)).numpy()
    np.savez_compressed(os.path.join(HERE, 'tiny_gen_grads.npz'), **out)
    print('tiny_gen_grads.npz', float(loss), len(out))


def gen_full():
    G = build_ref(FULL, seed=0)
    out = {}
    # checksums of the seeded random-init state (the 234 MB of weights are not committed)
    for k, v in G.state_dict().items():
        v = v.double().flatten()
        out['S.' + k] = np.asarray([v.sum().item(), v.abs().sum().item(), v[:: max(1, v.numel() // 7)].sum().item()])
    z, c, x, u8 = synth_inputs(FULL, 2, seed=11)
    taps = {}
    hook_taps(G, taps)
    with torch.no_grad():
        y = G(z, c, x, noise_mode='const')
    out.update({'z': z.numpy(), 'c': c.numpy(), 'x_u8': u8.numpy(), 'y': y.numpy()})
    for k, v in taps.items():                       # per-layer statistics + a small crop of every activation
        out['stat.' + k] = np.asarray([v.double().mean().item(), v.double().std().item(), v.abs().max().item()])
        if v.ndim == 4:
            out['crop.' + k] = v[:, :4, 5:13, 5:13].numpy().copy()
        else:
            out['crop.' + k] = v[:, :64].numpy().copy()
    # batch independence: sample 0 alone (style normalisation couples the batch at ~1e-6)
    with torch.no_grad():
        y0 = G(z[:1], c[:1], x[:1], noise_mode='const')
    out['y0_alone'] = y0.numpy()
    np.savez_compressed(os.path.join(HERE, 'full_gen.npz'), **out)
    print('full_gen.npz', y.shape, float(y.abs().max()), float((y0 - y[:1]).abs().max()))


if __name__ == '__main__':
    which = sys.argv[1:] or ['ops', 'tiny', 'tiny_grads', 'full']
    if 'ops' in which:
        gen_ops()
    if 'tiny' in which:
        gen_tiny()
    if 'tiny_grads' in which:
        gen_tiny_grads()
    if 'full' in which:
        gen_full()
