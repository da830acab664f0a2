import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """liboracle.so binding; built on demand (gcc only)."""
    import oracle_lib
    oracle_lib.build()
    return oracle_lib


@pytest.fixture(scope="session")
def gibbs():
    """The product binding.  GPU tests must fail loudly -- never skip -- when the library is missing."""
    from lda_thesis_b200 import _lib
    _lib.load_library()
    return _lib


def load_golden(name):
    import numpy as np
    with np.load(os.path.join(GOLDEN, name)) as f:
        return {k: f[k] for k in f.files}
