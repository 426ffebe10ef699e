import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "fs-eend_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA sm_100 (B200) device")


@pytest.fixture(scope="session")
def built_lib():
    """Path of the in-tree shared library (built on demand; nvcc cross-compiles without a GPU)."""
    from fseend_b200.build import build
    return build()
