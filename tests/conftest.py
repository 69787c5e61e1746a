import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _cuda_devices():
    try:
        import ctypes
        n = ctypes.c_int(0)
        rt = ctypes.CDLL("libcudart.so.12")
        return n.value if rt.cudaGetDeviceCount(ctypes.byref(n)) == 0 else 0
    except OSError:
        try:
            import torch
            return torch.cuda.device_count()
        except Exception:
            return 0


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a machine without a GPU: gpu-marked tests are skipped, not failed (the product
    itself still refuses to run without a device -- gtnb_ctx_create fails loudly, tests/test_capi_load.py)."""
    if not any("gpu" in it.keywords for it in items) or _cuda_devices() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


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
