import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    pyoracle.liboracle()
    return pyoracle


@pytest.fixture(scope="session")
def ctx():
    from gtn_b200 import capi
    c = capi.Ctx(0)
    yield c
    c.close()
