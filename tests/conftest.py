import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def native_lib():
    """Builds (if stale) and loads libsustaingym_b200.so."""
    from sustaingym_b200 import _build, _lib
    _build.build()
    return _lib.load()


GOLDEN = os.path.join(ROOT, "tests", "golden")
