import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _has_gpu() -> bool:
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
def lego():
    """The two frozen descriptor sets the reference ships (tests/golden/make_golden.py)."""
    left = np.load(os.path.join(GOLDEN, "lego_left.npz"))
    right = np.load(os.path.join(GOLDEN, "lego_right.npz"))
    return {
        "left": left["desc"], "right": right["desc"],
        "left_coord": left["coord"], "right_coord": right["coord"],
        "l2r": np.load(os.path.join(GOLDEN, "lego_l2r_expected.npy")),
        "r2l": np.load(os.path.join(GOLDEN, "lego_r2l_expected.npy")),
        "dist": np.load(os.path.join(GOLDEN, "lego_distances.npz")),
        "twin": np.load(os.path.join(GOLDEN, "lego_python_twin.npz"))["rows"],
    }


@pytest.fixture(scope="session")
def matcher():
    from photogrammetry_b200.keypoint_matching import Matcher
    m = Matcher(0)
    yield m
    m.close()
