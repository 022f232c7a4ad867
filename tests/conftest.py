import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a B200 (run with -m gpu on the GPU box)')


@pytest.fixture(scope='session', autouse=True)
def _built_library():
    """Make sure libxrt_b200.so exists (nvcc cross-compiles on the CPU box; on the GPU box it is prebuilt)."""
    import __graft_entry__ as g
    g.build()


@pytest.fixture(scope='session')
def cuda_device():
    from xrt_b200 import _lib
    return _lib.init(0)


def golden(name):
    import numpy as np
    return np.load(os.path.join(GOLDEN, name + '.npz'))
