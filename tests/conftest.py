"""pytest configuration: the ``gpu`` marker selects tests that need a real B200."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def production_configs():
    """Production filter parameters (/root/reference/code/run_capsule.py:377-388)."""
    no_cells = {"wavelet": "db3", "level": None, "sigma": 128, "max_threshold": 12}
    cells = {"wavelet": "db3", "level": None, "sigma": 64, "max_threshold": 3}
    return no_cells, cells
