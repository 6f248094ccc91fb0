import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def kamr_lib():
    """Loads libkamr.so (builds it when missing).  GPU tests go through this C-ABI only."""
    import __graft_entry__ as g
    from kitamr_jl_b200 import abi
    if not os.path.exists(abi.LIB_PATH):
        g.build()
    return abi.load()
