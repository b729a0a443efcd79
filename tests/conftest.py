import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run by the driver with -m gpu)")


def _has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without a GPU must fail loudly, not skip: the CUDA path has no fallback.
    # `-m "not gpu"` never touches these items.
    pass


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    z = np.load(os.path.join(ROOT, "tests", "golden", "csa_golden.npz"))
    cases = {}
    for k in z.files:
        name, field = k.split("/")
        cases.setdefault(name, {})[field] = z[k]
    return cases


@pytest.fixture(scope="session")
def mb():
    import mambo_b200
    return mambo_b200
