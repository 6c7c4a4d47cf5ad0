import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# Every test process fills the library's scratch buffers (partial sums, j-side rows) with NaN patterns when they are allocated
# (read once, at the first launch): a partial sum that is read without having been written then shows as a non-finite force
# instead of hiding behind the zeros of fresh device memory -- which is how round 2's multi-rank chunk-mask bug stayed invisible
# until allocations were reused on a 2-GPU box.
os.environ.setdefault("STEPS_B200_POISON", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """the library and the checkers must exist; building them is not using them"""
    from steps_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as ge

        ge.build()
    from oracle import pyport

    pyport.load()
