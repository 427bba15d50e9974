"""pytest configuration: `gpu` marker + shared fixtures.

CPU tests (`-m "not gpu"`) cover the oracle against the committed golden vectors, the host
logic and the C-ABI's symbol table.  GPU tests (`-m gpu`) are the parity tests proper and
call the CUDA path through the C-ABI (libb200tip.so).
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


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
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name), allow_pickle=False)

    return load


def case_names(npz, suffix):
    return sorted({k.rsplit(".", 1)[0] for k in npz.files if k.endswith("." + suffix)})
