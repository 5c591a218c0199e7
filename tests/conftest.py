import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle library (test infrastructure only)."""
    import oracle_lib
    return oracle_lib.load()


@pytest.fixture(scope="session")
def cuda():
    """The product library; GPU tests fail loudly if it is missing."""
    from clic_b200 import estimator
    return estimator.cuda_library()
