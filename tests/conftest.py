"""pytest configuration: the ``gpu`` marker selects tests that need a real B200."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionstart(session):
    """The CUDA library is a git-ignored build artefact: build it (nvcc cross-compiles without a
    GPU) when a fresh checkout runs the suite before `__graft_entry__.build()` was called."""
    lib = os.path.join(ROOT, "aind_smartspim_destripe_b200", "lib", "libdstr_b200.so")
    if not os.path.exists(lib):
        import __graft_entry__

        __graft_entry__.build()


@pytest.fixture(scope="session")
def production_configs():
    """Production filter parameters (/root/reference/code/run_capsule.py:377-388)."""
    no_cells = {"wavelet": "db3", "level": None, "sigma": 128, "max_threshold": 12}
    cells = {"wavelet": "db3", "level": None, "sigma": 64, "max_threshold": 3}
    return no_cells, cells
