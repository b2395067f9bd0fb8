import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    return load


def pytest_collection_modifyitems(config, items):
    # GPU tests are selected explicitly with -m gpu; without a CUDA device they are skipped, never faked.
    try:
        import torch
        has = torch.cuda.is_available()
    except Exception:
        has = False
    if not has:
        skip = pytest.mark.skip(reason="no CUDA device in this container")
        for it in items:
            if "gpu" in it.keywords:
                it.add_marker(skip)
