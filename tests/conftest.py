import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name))


def golden_state_dict(genre: bool, dtype=torch.float32):
    """The seeded reference-format state_dict of tests/golden/weights_seed0.npz."""
    w = load_golden("weights_seed0.npz")
    sd = {}
    for k in w.files:
        if not genre and k.startswith("genre_classifier."):
            continue
        t = torch.from_numpy(w[k])
        sd[k] = t.to(dtype) if t.is_floating_point() else t
    return sd


def float_state_dict(sd):
    return {k: v for k, v in sd.items() if not k.endswith("num_batches_tracked")}


@pytest.fixture(scope="session")
def fwd_golden():
    return load_golden("pcn_fwd.npz")


@pytest.fixture(scope="session")
def eq_golden():
    return load_golden("equivariance.npz")
