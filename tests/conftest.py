"""Shared pytest configuration: the `gpu` marker and golden-fixture helpers."""
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


def golden(name):
    """Load tests/golden/<name>.npz (generated from the unmodified reference)."""
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def trajectory_names():
    return sorted(f[len("traj_"):-4] for f in os.listdir(GOLDEN) if f.startswith("traj_"))


def traj_kwargs(g):
    """Constructor switches stored with a trajectory."""
    return {k[3:]: bool(g[k]) for k in g.files if k.startswith("kw_")}


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
