import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    # a gpu test on a machine without a GPU is skipped, never silently passed on a fallback
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _cuda_library_is_current():
    """(Re)build libharmonies_b200.so when it is missing or older than its sources (nvcc is in
    the image on both the dev container and the GPU box).  The product never builds or falls
    back on its own: without the library it raises."""
    try:
        from harmonies_alphazero_b200 import build

        build.build()
    except Exception as e:  # noqa: BLE001
        print(f"[conftest] could not build the CUDA library: {e}")
    yield


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as orc

    orc.build()
    return orc
