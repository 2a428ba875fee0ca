import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def port():
    """The plain-C oracle (built on demand; test infrastructure only)."""
    import oracle
    return oracle.Port()


@pytest.fixture(scope="session")
def lib():
    """The product library; the GPU tests must run through it."""
    import pytrimal_b200
    from pytrimal_b200 import build
    if not os.path.exists(pytrimal_b200.LIB_PATH):
        build.build_library()
    return pytrimal_b200.load()


@pytest.fixture(scope="session")
def gpu(lib):
    import pytrimal_b200
    if pytrimal_b200.device_count() < 1:
        pytest.fail("GPU test selected but libtrimal_cuda sees no sm_100 device "
                    "(there is no CPU fallback to test)")
    return pytrimal_b200


GOLDEN = os.path.join(ROOT, "tests", "golden")


def random_msa(rng, n, L, alphabet=b"ARNDCQEGHILKMFPSTWYV", gap=0.2, indet=0.02, lower=0.0,
               extra=b""):
    """Random rows over `alphabet` + gaps + 'X' (+ optional lower case / extra bytes)."""
    pool = np.frombuffer(alphabet + extra, np.uint8)
    m = pool[rng.integers(0, len(pool), (n, L))].copy()
    if lower:
        lo = rng.random((n, L)) < lower
        m[lo] |= 0x20
    m[rng.random((n, L)) < indet] = ord("X")
    m[rng.random((n, L)) < gap] = ord("-")
    return m
