import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference + probe (oracle/_ref).  Test infrastructure only."""
    from oracle import ref as oref
    if not oref.available():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    return oref.Ref()
