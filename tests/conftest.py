import ctypes
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def _no_device_reason():
    """None when fb_init(0) succeeds, else why gpu-marked tests cannot run here (the product has no CPU fallback)."""
    try:
        from floria_b200 import api

        lib = api.load_library()
    except Exception:  # library not built: no skip, the tests fail loudly (a missing extension is never hidden)
        return None
    h = ctypes.c_void_p()
    rc = lib.fb_init(0, ctypes.byref(h))
    if rc == 3:  # FB_ERR_NODEV: the only reason to skip
        return "no CUDA device: " + lib.fb_last_error(None).decode()
    if rc == 0:
        lib.fb_destroy(h)
    return None


def pytest_collection_modifyitems(config, items):
    gpu_items = [it for it in items if "gpu" in it.keywords]
    if not gpu_items:
        return
    why = _no_device_reason()
    if why is None:
        return
    if os.environ.get("FB_REQUIRE_GPU"):  # the GPU box: a missing device / library is a failure, not a skip
        raise pytest.UsageError("gpu tests requested but " + why)
    skip = pytest.mark.skip(reason=why)
    for it in gpu_items:
        it.add_marker(skip)
