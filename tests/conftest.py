import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests', 'golden')):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (runs on the B200 box)')


def pytest_collection_modifyitems(config, items):
    """gpu-marked tests need an sm_100 device: skip them (instead of failing in the driver probe) anywhere else."""
    try:
        import torch
        ok = torch.cuda.is_available() and torch.cuda.get_device_capability(0)[0] == 10
    except Exception:
        ok = False
    if ok:
        return
    skip = pytest.mark.skip(reason='needs a CUDA sm_100 device (B200)')
    for it in items:
        if 'gpu' in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope='session')
def golden_dir():
    return os.path.join(ROOT, 'tests', 'golden')


@pytest.fixture(scope='session')
def lib():
    """The product C-ABI library (built in-tree if missing)."""
    from lsnet_b200 import build as _b
    import ctypes
    path = _b.build()
    return ctypes.CDLL(path)
