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
def lib_path():
    """Build (if needed) and return the C-ABI library path.  nvcc cross-compiles without a GPU."""
    from regnet_for_3d_grasping_b200 import build
    return build.build()


@pytest.fixture(scope="session")
def oracle():
    from oracle import pn2_oracle
    pn2_oracle.build()
    return pn2_oracle


def golden(name):
    import numpy as np
    path = os.path.join(GOLDEN, name)
    if not os.path.exists(path):
        pytest.skip(f"golden fixture {name} not present")
    return np.load(path, allow_pickle=False)
