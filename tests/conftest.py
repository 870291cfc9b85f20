import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (oracle/libb2ref.so), built on demand.  Test infrastructure only."""
    from tests import oracle as o
    o.lib()
    return o


@pytest.fixture(scope="session")
def b2():
    """The product package; the CUDA library must have been built (no fallback)."""
    import dplug_b200
    from dplug_b200 import _lib
    _lib.lib()
    return dplug_b200
