import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    """True when the engine sees a CUDA device.  A box that HAS a GPU but cannot load the library (stale .so, ABI mismatch,
    missing build) must fail loudly, never skip: a silent skip would read as 'GPU tests green'."""
    try:
        from neuralode4j_b200 import _lib
        return _lib.load().node4j_device_count() > 0
    except Exception as e:
        if os.path.exists("/dev/nvidia0") or os.path.exists("/dev/nvidiactl"):
            raise pytest.UsageError(f"a CUDA device is present but libnode4j_b200.so cannot be used: {e!r}")
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
