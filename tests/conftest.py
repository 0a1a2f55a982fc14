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


def pytest_sessionfinish(session, exitstatus):
    """measured parity errors next to their bars (tests/helpers.py within()), for profiles/"""
    import json
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import helpers
    if not helpers.MEASURED:
        return
    out = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_measured.json"), "w") as f:
            json.dump(helpers.MEASURED, f, indent=1)
    except OSError:
        pass
