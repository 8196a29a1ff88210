import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    return dict(np.load(os.path.join(GOLDEN, "structures_golden.npz")))


@pytest.fixture(scope="session")
def built_lib():
    """Builds libtomoenc.so in-tree if needed (nvcc cross-compiles without a GPU)."""
    from tomoencoders_b200 import _lib
    _lib.build_library()
    return _lib


def seeded_volume(shape, dtype, seed):
    """Same generator as tests/golden/make_golden.py."""
    g = np.random.default_rng(seed)
    if np.dtype(dtype) == np.uint8:
        return g.integers(0, 256, shape, dtype=np.uint8)
    if np.dtype(dtype) == np.uint16:
        return g.integers(0, 65536, shape, dtype=np.uint16)
    return g.standard_normal(shape).astype(dtype)
