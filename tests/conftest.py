import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run on the GPU box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    return importlib.import_module("3d-point-cloud-multiday-imagery_b200")


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name))
    return load


@pytest.fixture(scope="session")
def engine(pkg):
    """One libmdkm handle on cuda:0.  Fails loudly (no fallback) when the extension or GPU is missing."""
    eng = pkg.Engine(0)
    yield eng
    eng.close()
