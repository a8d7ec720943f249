import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def oracle_ld():
    from oracle.oracle import Oracle
    return Oracle(long_double=True)


@pytest.fixture(scope="session")
def gold_cg():
    return dict(np.load(os.path.join(GOLD, "celdaCG.npz")))


@pytest.fixture(scope="session")
def gold_c():
    return dict(np.load(os.path.join(GOLD, "celdaC.npz")))


@pytest.fixture(scope="session")
def gold_g():
    return dict(np.load(os.path.join(GOLD, "celdaG.npz")))


@pytest.fixture(scope="session")
def gold_grid():
    return dict(np.load(os.path.join(GOLD, "celdaCGGrid.npz")))
