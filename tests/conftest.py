import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built():
    """Builds (if stale) the CUDA library and the oracle; both builds work without a GPU."""
    import __graft_entry__ as g

    g.build()
    return True


@pytest.fixture(scope="session")
def engine(built):
    import dual_threshold_optimization_b200 as dto

    eng = dto.Engine(0)  # raises DtoError without a GPU: the gpu-marked tests must not silently pass
    yield eng
    eng.close()
