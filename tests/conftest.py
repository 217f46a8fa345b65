import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def _have_gpu():
    """Asks the library itself (rsv_device_count: the CUDA runtime it links), not torch: a GPU box without torch must
    still run the parity tests. RSV_REQUIRE_GPU=1 turns "no device" into a failure instead of a skip."""
    try:
        from resolv_b200 import _abi
        return _abi.load().rsv_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    if os.environ.get("RSV_REQUIRE_GPU", "0") not in ("", "0"):
        raise pytest.UsageError("RSV_REQUIRE_GPU is set but libresolv_b200.so reports no CUDA device (or is not built)")
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
