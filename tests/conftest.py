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


def pytest_collection_modifyitems(config, items):
    """A GPU test that hangs (a kernel that never returns) must end the PROCESS, not sit on the
    device until the box is reclaimed: hard per-test limit through pytest-timeout's thread method."""
    if not config.pluginmanager.hasplugin("timeout"):
        return
    for item in items:
        if item.get_closest_marker("gpu") and not item.get_closest_marker("timeout"):
            item.add_marker(pytest.mark.timeout(600, method="thread"))


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
