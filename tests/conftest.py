import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def built_lib():
    """libmmx.so built in-tree (cross-compiles on the CPU box)."""
    import mmx_b200
    if not os.path.exists(mmx_b200.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return mmx_b200.lib()
