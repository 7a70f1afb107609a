import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """GPU run: the kernel parity tests first (seconds), the command-line replays (one process per call) last."""
    items.sort(key=lambda it: 1 if "test_gpu_cli" in it.nodeid else 0)
