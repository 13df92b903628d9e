import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "ref: needs oracle/_ref (the compiled unmodified reference)")


@pytest.fixture(scope="session")
def port():
    import oracle
    oracle.build()
    return oracle.port()


@pytest.fixture(scope="session")
def ref_fm():
    import oracle
    oracle.build()
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    return oracle.RefFm()


@pytest.fixture(scope="session")
def ref_power():
    import oracle
    oracle.build()
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    return oracle.RefPower()
