import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import orc
    return orc.Oracle()


@pytest.fixture(scope="session")
def refsam():
    from oracle import orc
    orc.build()
    if not orc.ref_sam_available():
        pytest.skip("oracle/_ref/libsaltref_sam.so not built (reference tree absent)")
    return orc.RefSam()


@pytest.fixture(scope="session")
def ref():
    from oracle import orc
    if not orc.ref_available():
        pytest.skip("oracle/_ref not built (reference tree absent)")
    return orc.Ref()
