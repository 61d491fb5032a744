import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """Plain ``pytest tests`` on a machine without CUDA: skip the gpu-marked tests instead of failing."""
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="needs CUDA (B200); run with -m gpu on the GPU box")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def built_lib():
    """The in-tree libsaa_b200.so (built on demand; nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as g
    g.build()
    from riskaversetrajopt_b200 import _lib
    return _lib


@pytest.fixture(scope="session")
def drone_seed0():
    """Reference inputs: np.random.seed(0), first sample_uncertain_parameters('saa', M=50)
    (drone/drone_risk.py:57, :483-484)."""
    from riskaversetrajopt_b200.drone.drone_utils import sample_uncertain_parameters
    state = np.random.get_state()
    np.random.seed(0)
    out = sample_uncertain_parameters('saa', M=50)
    np.random.set_state(state)
    return out


def rel_err(a, b, floor=1e-12):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))) if a.size else 0.0
