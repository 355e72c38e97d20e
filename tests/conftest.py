import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line(
        'markers', 'gpu: test needs a CUDA device (B200); run with -m gpu')


@pytest.fixture(scope='session')
def native_lib():
    """Makes sure the native library exists (builds it if nvcc is here)."""
    from chi_b200 import _build, _lib
    if not os.path.exists(_lib.LIB_PATH):
        _build.build_core()
    return _lib.lib()
