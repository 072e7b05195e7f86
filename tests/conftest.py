import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built_lib():
    """Build (or reuse) libnerfb200.so; nvcc cross-compiles without a GPU."""
    from nerf_rep_for_test_b200 import build
    return build.build()


def golden(name):
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))


@pytest.fixture(autouse=True)
def _seed_torch():
    """Every test starts from the same torch RNG state (CPU and CUDA): the stratified jitter and the random u of the
    training-mode renders draw from the default generators, so without this a loss-decrease threshold would see a
    different sample set in every run.  (What stays run-to-run variable on the GPU is the order of the split-K atomics.)"""
    import torch
    torch.manual_seed(20240229)
    yield
