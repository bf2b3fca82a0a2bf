import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "needs_reference: needs the read-only reference tree at /root/reference")


def pytest_collection_modifyitems(config, items):
    from oracle import live_reference as live
    if live.available():
        return
    skip = pytest.mark.skip(reason="reference tree not present on this machine")
    for item in items:
        if "needs_reference" in item.keywords:
            item.add_marker(skip)
