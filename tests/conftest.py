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
def gpu(built, request):
    """The product library on a B200.  Under `-m gpu` (the GPU box) a missing device is a hard failure: GPU tests must never
    pass on a fallback.  In an unfiltered run on a machine without a device the GPU tests are skipped instead."""
    import gravit_b200 as g

    try:
        g.init(0)
    except g.EngineError:
        expr = request.config.getoption("-m") or ""
        if "gpu" in expr and "not gpu" not in expr:
            raise
        pytest.skip("no CUDA device (run with -m gpu on a B200 box)")
    return g
