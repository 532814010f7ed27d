import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    """CPU oracle (oracle/liboracle.so) -- the checker, never the product."""
    from oracle import pyoracle

    pyoracle.cpu()
    return pyoracle


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    return np.load(os.path.join(ROOT, "tests", "golden", "golden_mt.npz"))


@pytest.fixture(scope="session")
def b2s():
    """The product C-ABI library, loaded through ctypes (fails loudly if it is not built)."""
    from cub_b200 import _lib

    return _lib.load()


@pytest.fixture(scope="session")
def refcub():
    """Unmodified reference CUB 2.2.0 built from /root/reference into oracle/_ref (GPU)."""
    from oracle import pyoracle

    lib = pyoracle.load_gpu_reference("ref")
    if lib is None:
        pytest.skip("oracle/_ref/libref_cub.so not built (needs /root/reference at build time)")
    return lib
