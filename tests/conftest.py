import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


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
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    import oracle as O  # oracle/oracle.py — test infrastructure

    O.lib()
    return O


@pytest.fixture(scope="session")
def hostcheck():
    """g++ build of the kernels' per-thread math (tests/hostcheck/hostcheck.cpp)."""
    import ctypes as C

    d = os.path.join(ROOT, "tests", "hostcheck")
    so = os.path.join(d, "libhostcheck.so")
    src = os.path.join(d, "hostcheck.cpp")
    hdr = os.path.join(ROOT, "flgp_b200", "csrc", "core_math.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-mfma", "-ffp-contract=off", "-fPIC", "-shared", "-x",
                               "c++", src, "-o", so])
    return C.CDLL(so)


@pytest.fixture(scope="session")
def flgp():
    """The product API; importing it loads libflgp_b200.so (built by __graft_entry__.build())."""
    import flgp_b200

    flgp_b200._lib.load()
    return flgp_b200


@pytest.fixture(scope="session")
def ctx(flgp):
    return flgp.default_ctx()


def spiral(n, seed=0):
    rng = np.random.default_rng(seed)
    th = rng.uniform(0, 8 * np.pi, n)
    X = np.c_[(th + 4) ** 0.7 * np.cos(th), (th + 4) ** 0.7 * np.sin(th)]
    Y = 3 * np.sin(th / 10) + 3 * np.cos(th / 2) + 4 * np.sin(4 * th / 5) + rng.standard_normal(n)
    return np.asfortranarray(X), Y


def swiss(n, seed=0):
    rng = np.random.default_rng(seed)
    t = rng.uniform(1.5 * np.pi, 4.5 * np.pi, n)
    h = rng.uniform(0, 21, n)
    X = np.c_[t * np.cos(t), h, t * np.sin(t)]
    Y = np.sin(t) + h / 21 + 0.1 * rng.standard_normal(n)
    return np.asfortranarray(X), Y
