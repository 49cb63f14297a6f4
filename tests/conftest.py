import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def built_lib():
    """Builds (if stale) and loads libcmf_b200.so; CPU-only boxes can do this too."""
    import __graft_entry__ as g
    g.build()
    from cmfpy_b200 import _lib
    return _lib.load()


def golden(name):
    import numpy as np
    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    if not os.path.exists(path):
        pytest.skip("golden fixture %s not generated" % name)
    return np.load(path)
