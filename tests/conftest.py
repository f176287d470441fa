import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def port():
    from oracle import loader
    return loader.port()


@pytest.fixture(scope="session")
def ref():
    """Compiled reference (oracle/_ref/libref_oai.so); tests that need it skip when it
    is neither prebuilt nor buildable (no /root/reference)."""
    from oracle import loader
    r = loader.ref()
    if r is None:
        pytest.skip("compiled reference oracle/_ref/libref_oai.so not available")
    return r
