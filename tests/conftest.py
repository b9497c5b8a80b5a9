import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
RUN_SEED = 20261019


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


def _has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def classic():
    from oponn_b200.maps import classic42
    return classic42(6)


@pytest.fixture(scope="session")
def rules():
    from oponn_b200.state import Rules
    return Rules()


@pytest.fixture(scope="session")
def oracle(classic, rules):
    from oracle.pyoracle import Oracle
    o = Oracle(classic, rules)
    yield o
    o.close()


@pytest.fixture(scope="session")
def engine(classic, rules):
    from oponn_b200.engine import RolloutEngine
    e = RolloutEngine(classic, rules, device=0)
    yield e
    e.close()


@pytest.fixture(scope="session")
def mid_leaf():
    return np.fromfile(os.path.join(GOLDEN, "classic42_mid_s1.leaf"), dtype=np.uint8)


@pytest.fixture(scope="session")
def golden_games():
    return np.load(os.path.join(GOLDEN, "golden_games.npz"))
