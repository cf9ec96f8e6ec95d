"""Shared test plumbing.

`-m "not gpu"` : oracle vs the golden vectors of the real reference, host logic, ABI symbol checks,
                 world_size-2 gloo tests (all CPU, a few minutes).
`-m gpu`       : the parity tests proper -- CUDA path (through the C ABI) vs the oracle.
Nothing here reads /root/reference at run time (it does not exist on the GPU box).
"""
import importlib
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
    config.addinivalue_line("markers", "gpu: needs a CUDA (sm_100) device; run with -m gpu on the B200 box")


def load_pkg():
    name = "4dcapture-fpv_b200"
    if name not in sys.modules:
        if not os.path.exists(os.path.join(ROOT, name, "libfpv_b200.so")):
            sys.path.insert(0, os.path.join(ROOT, name))
            import build as fpv_build  # 4dcapture-fpv_b200/build.py
            fpv_build.build_library()
            sys.path.pop(0)
        mod = importlib.import_module(name)
        sys.modules["fpv_b200"] = mod
    return sys.modules[name]


@pytest.fixture(scope="session")
def fpv():
    return load_pkg()


@pytest.fixture(scope="session")
def cuda_dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def golden_cases():
    """Chamfer goldens (tests/golden/make_golden.py); the prior_*.npz files belong to test_*prior*.py."""
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz") and not f.startswith("prior_"))


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))
