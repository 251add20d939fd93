import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def pytest_collection_modifyitems(config, items):
    """Tests marked `gpu` are skipped (not failed) on a machine without a CUDA device, so a plain `pytest` is green on the
    CPU tier as well as with `-m "not gpu"`."""
    try:
        import torch
        have_cuda = torch.cuda.is_available()
    except Exception:
        have_cuda = False
    if have_cuda:
        return
    skip = pytest.mark.skip(reason='no CUDA device (GPU tier: run with -m gpu on the B200 box)')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session')
def golden_ops():
    import numpy as np
    return np.load(os.path.join(GOLDEN, 'ops.npz'))


@pytest.fixture(scope='session')
def golden_tiny():
    import numpy as np
    return np.load(os.path.join(GOLDEN, 'tiny_gen.npz'))


@pytest.fixture(scope='session')
def golden_full():
    import numpy as np
    return np.load(os.path.join(GOLDEN, 'full_gen.npz'))


def rel_err(a, b):
    """max|a-b| / max|b| -- the tolerance metric used throughout."""
    import numpy as np
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.fixture(scope='session')
def golden_tiny_grads():
    import numpy as np
    return np.load(os.path.join(GOLDEN, 'tiny_gen_grads.npz'))
