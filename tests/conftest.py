import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run with -m gpu on a B200)')


def _gpu_unavailable():
    """Reason string when the `gpu` tests cannot run here (no device or no built library)."""
    try:
        import torch
        if not torch.cuda.is_available():
            return 'no CUDA device'
    except Exception as e:                               # noqa: BLE001
        return 'torch unavailable: %s' % e
    try:
        from deconv3d_b200 import _native
        _native.load()
    except Exception as e:                               # noqa: BLE001
        return 'libdeconv3d_b200.so does not load: %s' % e
    return None


def pytest_collection_modifyitems(config, items):
    # a plain `pytest` on a CPU box skips the GPU tests instead of failing them; with `-m gpu`
    # (the GPU tier) a missing device or library is an ERROR, never a silent skip
    marker_expr = config.getoption('-m') or ''
    if 'gpu' in marker_expr and 'not gpu' not in marker_expr:
        return
    why = _gpu_unavailable()
    if why is None:
        return
    skip = pytest.mark.skip(reason='gpu test: ' + why)
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False) as f:
        return {k: f[k] for k in f.files}


@pytest.fixture(scope='session')
def golden():
    return load_golden
