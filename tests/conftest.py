import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built_lib():
    from squid_b200 import build
    return build.build(verbose=False)


@pytest.fixture(scope="session")
def ref_oracle():
    from oracle import pyref
    pyref.build()
    if not pyref.available():
        pytest.skip("oracle/_ref/squid_ref not built (needs /root/reference once)")
    return pyref
