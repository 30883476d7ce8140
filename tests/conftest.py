import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 via gpurun)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the in-tree libraries once if they are missing (CPU box: nvcc cross-compiles)."""
    import __graft_entry__ as g

    g.build(only_if_missing=True)
