import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
# kernels specialised at run time (ec_set_lazy(3)) keep their cubins on disk: tests use a scratch directory, not ~/.cache
import tempfile  # noqa: E402

os.environ.setdefault("EC_JIT_CACHE", os.path.join(tempfile.gettempdir(), "erased_cells_b200_jit_tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionstart(session):
    """The product is a compiled library: build it in-tree if this checkout has not been built yet
    (nvcc cross-compiles for sm_100a without a GPU; __graft_entry__.build() does the same)."""
    from erased_cells_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure): compiled on demand with g++."""
    from oracle import oracle

    oracle.build()
    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def landsat():
    import numpy as np

    z = np.load(os.path.join(ROOT, "tests", "golden", "landsat_elkton.npz"))
    return {k: z[k] for k in z.files}
