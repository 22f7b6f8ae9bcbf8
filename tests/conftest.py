import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _has_gpu():
    try:
        import ctypes as C
        from posepipeline_b200 import _lib
        n = C.c_int()
        return _lib.load().pe_device_count(C.byref(n)) == 0 and n.value > 0
    except Exception:
        return False


HAS_GPU = _has_gpu()


def pytest_collection_modifyitems(config, items):
    for item in items:
        if "gpu" in item.keywords and not HAS_GPU:
            item.add_marker(pytest.mark.skip(reason="no CUDA device"))
