import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def orc():
    from oracle import sp_oracle
    sp_oracle.build()
    return sp_oracle


@pytest.fixture(scope="session")
def spb():
    import saddlepoint_b200
    saddlepoint_b200.capi.load()
    return saddlepoint_b200


@pytest.fixture(scope="session")
def handle(spb):
    h = spb.Handle(0)
    yield h
    h.close()
