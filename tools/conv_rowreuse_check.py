"""Development check of the row-reuse mode of the tcgen05 convolution (afcm_conv_tc_set_rowreuse): bit-exactness on
fp16-representable integer inputs, then timings at the AFCM layer shapes."""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from afcm_b200 import _lib
from afcm_b200.torch_utils.ops import conv2d_gradfix

dev = torch.device('cuda:0')
L = _lib.lib()
g = torch.Generator(device='cpu').manual_seed(3)


def exact(N, Ci, Co, H, W):
    x = torch.randint(-4, 5, (N, Ci, H, W), generator=g).float().to(dev)
    w = torch.randint(-2, 3, (Co, Ci, 3, 3), generator=g).float().to(dev)
    ref = conv2d_gradfix.conv2d_native(x, w, 2, impl='f32')
    got = conv2d_gradfix.conv2d_native(x, w, 2, impl='tc')
    torch.cuda.synchronize()
    return bool(torch.equal(ref, got)), float((ref - got).abs().max())


def timeit(N, Ci, Co, H):
    x = torch.randn(N, Ci, H, H, device=dev).half()
    w = torch.randn(Co, Ci, 3, 3, device=dev)
    ent = conv2d_gradfix.prepare_weight(w, 1.0, False, want_tc=True)
    plane = int(L.afcm_conv_tc_plane_elems(H, H, Ci))
    xp = torch.empty(N, plane, dtype=torch.float16, device=dev)
    y = torch.empty(N, Co, H + 2, H + 2, device=dev, dtype=torch.float16)
    st = _lib.stream_ptr(dev)
    _lib.check(L.afcm_conv_tc_pack(_lib.ptr(x), 1, None, _lib.ptr(xp), 1, N, Ci, H, H, st))
    fn = lambda: _lib.check(L.afcm_conv2d_tc(_lib.ptr(xp), _lib.ptr(ent[('w_tc', torch.float16)]), None, None, _lib.ptr(y), 1, 1,
                                             N, Ci, H, H, Co, 2, st))
    for _ in range(2):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        fn()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    return ms, 2.0 * N * Co * Ci * 9 * (H + 2) ** 2 / ms / 1e9


for mode, bo in [(0, 0), (1, 0), (4, 0)]:
    L.afcm_conv_tc_set_rowreuse(mode)
    try:
        r = [exact(2, 64, 64, 36, 36), exact(1, 192, 96, 30, 22), exact(2, 91, 128, 52, 52)]
    except Exception as e:
        r = repr(e)
    print('rowreuse', mode, 'base_ofs', bo, 'exact:', r, flush=True)

shapes = [(16, 362, 512, 148), (16, 181, 256, 148), (16, 4, 64, 276), (16, 64, 64, 276), (16, 64, 91, 276), (16, 91, 128, 276), (16, 128, 181, 276), (16, 181, 128, 148),
          (16, 128, 91, 276), (16, 91, 64, 276), (16, 256, 181, 148), (16, 512, 512, 84), (16, 512, 512, 36)]
good = int(os.environ.get('ROWREUSE_BO', '0'))
for sh in shapes:
    out = []
    for mode in (0, 1, 4):
        L.afcm_conv_tc_set_rowreuse(mode)
        try:
            out.append('%.3f ms %6.0f TF' % timeit(*sh))
        except RuntimeError as e:
            out.append('n/a')
    print(sh, ' per-tap:', out[0], ' row-reuse (combined stages):', out[1], ' row-reuse (split rings):', out[2], flush=True)

L.afcm_conv_tc_set_rowreuse(-1)
for mode in (0, 16, 48):
    L.afcm_conv_tc_debug_buffer(mode << 8)
    print('dbg mode', mode, [('%.3f ms' % timeit(*sh)[0]) for sh in [(16, 64, 64, 276), (16, 91, 128, 276), (16, 512, 512, 84)]], flush=True)
L.afcm_conv_tc_debug_buffer(0)
