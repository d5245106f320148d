"""pytest configuration: `gpu` marker, shared helpers.  CPU tests (`-m "not gpu"`) cover the oracle against the golden
vectors, the host logic and the C-ABI surface; GPU tests (`-m gpu`) are the parity tests proper and call through libovk."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) GPU; run with -m gpu on the GPU box")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name: str):
    path = os.path.join(GOLDEN_DIR, name)
    assert os.path.exists(path), f"golden fixture {name} missing (python -m oracle.make_golden regenerates it)"
    return dict(np.load(path))


@pytest.fixture(scope="session")
def golden():
    return load_golden
