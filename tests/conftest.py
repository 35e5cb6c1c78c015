import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def ob():
    """oracle bindings (checker)."""
    from oracle import bindings
    bindings.oracle()
    return bindings


@pytest.fixture(scope="session")
def yn():
    """the product's numpy front-end; loading fails loudly if the library is missing."""
    import yael_b200
    yael_b200.lib()
    from yael_b200 import ynumpy
    return ynumpy
