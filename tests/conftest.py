import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run on the GPU box with `-m gpu`)")


@pytest.fixture(scope="session")
def oracle():
    """CPU oracle (test infrastructure).  Built by `make -C oracle`."""
    from tests import oracle_lib
    return oracle_lib


@pytest.fixture(scope="session")
def gpu():
    """Initialised CUDA library; fails loudly (no CPU fallback) when there is no device."""
    import lighthouse_b200
    lighthouse_b200.init(0)
    return lighthouse_b200
