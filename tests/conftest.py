import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run by the driver on the GPU box with -m gpu)")


@pytest.fixture(scope="session")
def handle():
    """One libswne_nmf handle on cuda:0 for the whole GPU session.  No skip: a missing GPU or a missing
    extension must fail loudly (there is no CPU fallback)."""
    from swne_b200 import NMFHandle
    h = NMFHandle(0)
    yield h
    h.close()
