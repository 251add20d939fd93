"""CPU tier: the C-ABI library builds, loads without a GPU driver and exports every symbol that
include/afcm_b200.h declares; argument validation that needs no device works."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


@pytest.fixture(scope='module')
def lib():
    from afcm_b200 import build
    build.build()
    from afcm_b200 import _lib
    return _lib.lib()


def _declared():
    src = open(os.path.join(ROOT, 'include', 'afcm_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(afcm_[a-z0-9_]+)\s*\(', src)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    from afcm_b200 import _lib
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f'{n} declared in afcm_b200.h but not exported'
        assert n in _lib.SIGNATURES, f'{n} has no ctypes signature in afcm_b200/_lib.py'
    assert set(_lib.SIGNATURES) == set(names)


def test_no_link_dependency_on_the_driver():
    import subprocess
    from afcm_b200 import _lib
    out = subprocess.run(['ldd', _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert 'libcuda.so' not in out and 'libcudart' not in out and 'torch' not in out


def test_version_and_size_helpers(lib):
    assert lib.afcm_version() == 1
    yh, yw = ctypes.c_int(), ctypes.c_int()
    # enc0 geometry of SURVEY.md 8.0: 278 -> 276 with up 2 / down 2, 12 taps each, padding 9,8,9,8
    assert lib.afcm_filtered_lrelu_out_size(278, 278, 2, 2, 12, 12, 9, 8, 9, 8, yh, yw) == 0
    assert (yh.value, yw.value) == (276, 276)
    assert lib.afcm_filtered_lrelu_out_size(278, 278, 2, 4, 12, 24, 34, 33, 34, 33, yh, yw) == 0
    assert (yh.value, yw.value) == (148, 148)
    assert lib.afcm_filtered_lrelu_out_size(38, 38, 4, 2, 24, 12, -6, -9, -6, -9, yh, yw) == 0
    assert (yh.value, yw.value) == (52, 52)
    sh, swb = ctypes.c_int(), ctypes.c_int()
    lib.afcm_filtered_lrelu_sign_size(276, 276, 2, 12, sh, swb)
    assert (sh.value, swb.value) == (276 * 2 - 1 + 11, ((276 * 2 - 1 + 11 + 15) // 16 * 16) // 4)
    assert lib.afcm_conv_tc_plane_elems(36, 36, 362) == 36 * 38 * 368
    # invalid arguments are reported through the status code + message, never by crashing
    assert lib.afcm_filtered_lrelu_out_size(4, 4, 2, 2, 12, 12, 0, 0, 0, 0, yh, yw) == -2
    assert b'upsampled buffer' in lib.afcm_last_error()


def test_ops_refuse_cpu_tensors():
    import torch
    from afcm_b200.torch_utils.ops import filtered_lrelu, bias_act, upfirdn2d
    x = torch.zeros(1, 1, 8, 8)
    for fn in (lambda: filtered_lrelu.filtered_lrelu(x), lambda: bias_act.bias_act(x), lambda: upfirdn2d.upfirdn2d(x, None)):
        with pytest.raises(RuntimeError, match='no CPU fallback'):
            fn()


def test_setup_filter_matches_reference_convention():
    import torch
    from afcm_b200.torch_utils.ops import upfirdn2d
    f = upfirdn2d.setup_filter([1, 3, 3, 1])
    assert f.shape == (4, 4) and abs(float(f.sum()) - 1) < 1e-6
    f = upfirdn2d.setup_filter(list(range(1, 9)))
    assert f.shape == (8,)
    assert torch.allclose(f, torch.arange(1., 9.) / 36)


def test_host_generator_mirrors_reference_init(golden_full):
    """Same seeded random init and same registered filters as the reference generator."""
    import numpy as np
    from afcm_b200.networks_stylegan3 import afcm_generator
    G = afcm_generator(seed=0, device=None)
    sd = G.state_dict()
    g = golden_full
    keys = [k[2:] for k in g.files if k.startswith('S.')]
    assert set(keys) == set(sd.keys())
    for k in keys:
        v = sd[k].double().flatten()
        got = np.asarray([v.sum().item(), v.abs().sum().item(), v[:: max(1, v.numel() // 7)].sum().item()])
        assert np.allclose(got, g['S.' + k], rtol=0, atol=0), k
