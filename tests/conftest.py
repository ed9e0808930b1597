import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def built():
    """Make sure the native pieces exist (no-op when already built)."""
    import __graft_entry__ as ge

    ge.build()
    return True


@pytest.fixture(scope="session")
def gpu(built):
    import gravit_b200 as g

    g.init(0)  # raises without a device: GPU tests must never pass on a fallback
    return g
