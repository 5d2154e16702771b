import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ref():
    """the reference's own search core compiled as a checker (oracle/_ref/libcatnet_ref.so)"""
    from oracle import ref as r
    if not r.available():
        pytest.skip("oracle/_ref/libcatnet_ref.so not built (needs /root/reference at build time)")
    return r
