import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def drt_lib():
    """Build (if stale) and load libdartray_gpu.so."""
    from dartray_b200 import build, capi
    build.build()
    return capi.load()


@pytest.fixture(scope="session")
def oracle_mod():
    from tests import oracle_lib
    oracle_lib.build_oracle()
    return oracle_lib
