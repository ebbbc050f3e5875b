import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def pkg():
    from __graft_entry__ import load_package
    p = load_package()
    p.lib()
    return p


@pytest.fixture(scope="session")
def ref():
    import oracle_util
    if not oracle_util.have_ref("e0"):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    return oracle_util.RefOracle("e0")


@pytest.fixture(scope="session")
def ref_e1():
    import oracle_util
    if not oracle_util.have_ref("e1"):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    return oracle_util.RefOracle("e1")
