import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run on the GPU box with -m gpu)")
    config.addinivalue_line("markers", "timeout: per-test limit (pytest-timeout; inert when the plugin is absent)")


@pytest.fixture(scope="session")
def port():
    import oracle
    return oracle.port()


@pytest.fixture(scope="session")
def ref():
    import oracle
    r = oracle.ref()
    if r is None:
        pytest.skip("oracle/_ref/libsdrref.so not built (no /root/reference on this box)")
    return r
