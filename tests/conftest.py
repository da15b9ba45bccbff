import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def built_library():
    """libpapr_b200.so, built on the spot (nvcc cross-compiles without a GPU) when a fresh checkout has none yet."""
    from papr_b200 import build as B
    if not os.path.exists(B.LIB):
        if not os.path.exists(B.NVCC):
            pytest.skip("libpapr_b200.so is not built and nvcc is not available")
        B.build_library()
    return B.LIB
