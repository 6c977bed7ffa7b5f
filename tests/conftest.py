import importlib.util
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def load_pkg():
    """Import the product package (its directory name has a hyphen, like every ComfyUI custom node)."""
    import __graft_entry__ as ge
    return ge.load_package()


@pytest.fixture(scope="session")
def pkg():
    return load_pkg()
