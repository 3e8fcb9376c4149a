import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with `-m gpu` under gpurun)")


@pytest.fixture(scope="session")
def bb():
    import bosip_b200
    return bosip_b200


@pytest.fixture(scope="session")
def ctx(bb):
    """One CUDA context for the whole GPU session; fails loudly when the library or the GPU is missing."""
    return bb.Context(0)
