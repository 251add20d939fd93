"""Repeated direct-NCHW convolutions against the packed path at shapes of every pipeline mode: counts the launches that are not
bit-identical (how the raw-slot race of DESIGN.md 3.1 was isolated).  usage: [ISSUERS=1,2] python tools/conv_direct_repeat.py [reps]"""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from afcm_b200 import _lib
from afcm_b200.torch_utils.ops import conv2d_gradfix
dev = torch.device('cuda:0')
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 12
for shape in [(1, 64, 64, 276, 276), (1, 128, 96, 276, 276), (1, 128, 128, 276, 276), (1, 181, 181, 276, 276), (1, 256, 256, 148, 148), (2, 512, 512, 84, 84)]:
    N, Ci, Co, H, W = shape
    g = torch.Generator(device='cpu').manual_seed(31)
    x = torch.randn(N, Ci, H, W, generator=g).to(dev).half()
    w = torch.randn(Co, Ci, 3, 3, generator=g).to(dev)
    bias = torch.randn(Co, generator=g).to(dev)
    scale = 1.0 / np.sqrt(Ci * 9)
    conv2d_gradfix.set_conv_impl('f32', torch.float16)
    run = lambda: conv2d_gradfix.conv2d_native(x, w, 2, pre_scale=scale, impl='tc', out_dtype=torch.float16, bias=bias)
    conv2d_gradfix.direct_nchw = False
    base = run().cpu()
    for iss in [int(v) for v in os.environ.get('ISSUERS', '1,2').split(',')]:
        _lib.lib().afcm_conv_tc_set_issuers(iss)
        for direct in (False, True):
            conv2d_gradfix.direct_nchw = direct
            bad = []
            for rep in range(reps):
                v = run().cpu()
                ne = (v != base)
                if ne.any():
                    idx = ne.nonzero()
                    bad.append((rep, int(ne.sum()) // Co, idx[0].tolist()[2:], round(float((v.float() - base.float()).abs().max()), 3)))
            print(shape, 'issuers', iss, 'direct', direct, 'bad %d/%d' % (len(bad), reps), bad[:4], flush=True)
