import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def cuda_lib():
    """The C-ABI library, built if necessary.  GPU tests fail (not skip) if it cannot be loaded."""
    from casapose_b200 import _lib

    if _lib._sources_newer_than_lib():
        _lib.build()
    return _lib.lib()
