import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _device_count():
    try:
        from genmap_b200 import _lib
        return int(_lib.lib().gmb_device_count())
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    """GPU run: the kernel parity tests first (seconds), the command-line replays (one process per call) last.
    Without a CUDA device the `gpu` tests are skipped, not failed (a plain `pytest tests` on a CPU box stays green);
    GMB_REQUIRE_GPU=1 turns the skip back into a failure for runs that must not pass without the device."""
    items.sort(key=lambda it: 1 if "test_gpu_cli" in it.nodeid else 0)
    if any("gpu" in it.keywords for it in items) and _device_count() == 0 and os.environ.get("GMB_REQUIRE_GPU") != "1":
        skip = pytest.mark.skip(reason="no CUDA device: the map path has no CPU fallback")
        for it in items:
            if "gpu" in it.keywords:
                it.add_marker(skip)
