import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def built_library():
    """The in-tree C-ABI library, built here when the checkout is fresh (nvcc cross-compiles without a GPU); loading it
    and listing its symbols is all the CPU suite does with it."""
    from gnnlm_b200 import build
    return build.build()


@pytest.fixture(scope="session")
def dev():
    """cuda:0 for the `-m gpu` tests; fails loudly when the device or the in-tree extension is missing (no fallback)."""
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from gnnlm_b200 import _lib
    _lib.load()
    return torch.device("cuda:0")
