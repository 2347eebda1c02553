import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure libpxf.so and the oracle exist (built in-tree; both travel to the GPU box)."""
    so = os.path.join(ROOT, "pyxfocus_b200", "libpxf.so")
    oso = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
    if not (os.path.exists(so) and os.path.exists(oso)):
        import __graft_entry__ as g
        g.build()
    yield


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        return np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    return load
