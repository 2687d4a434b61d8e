import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def goldens():
    return json.load(open(os.path.join(GOLDEN, "ref_goldens.json")))


@pytest.fixture(scope="session")
def oracle_mod():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as mod   # oracle/oracle.py -- test infrastructure
    mod.load()
    return mod


@pytest.fixture(scope="session")
def lb():
    import lulesh_b200
    return lulesh_b200


def load_npz(name):
    return dict(np.load(os.path.join(GOLDEN, name)))
