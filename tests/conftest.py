import json
import os
import subprocess
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
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def audio_case(name, n=480000):
    """Same generator as scripts/make_golden.py."""
    rng = np.random.default_rng(sum(map(ord, name)))
    t = np.arange(n) / 16000.0
    if name == "noise":
        return (0.1 * rng.standard_normal(n)).astype(np.float32)
    if name == "sweep":
        return (0.5 * np.sin(2 * np.pi * (100.0 + 120.0 * t) * t)).astype(np.float32)
    if name == "zeros":
        return np.zeros(n, np.float32)
    if name == "impulse0":
        a = np.zeros(n, np.float32); a[0] = 1.0; return a
    if name == "impulseN":
        a = np.zeros(n, np.float32); a[-1] = 1.0; return a
    if name == "speechlike":
        am = 0.6 + 0.4 * np.sin(2 * np.pi * 3.0 * t)
        return (0.1 * am * rng.standard_normal(n) + 0.05 * np.sin(2 * np.pi * 220 * t)).astype(np.float32)
    raise KeyError(name)


@pytest.fixture(scope="session")
def golden_logmel():
    return np.load(os.path.join(GOLDEN, "logmel_golden.npz"))


@pytest.fixture(scope="session")
def golden_small():
    return np.load(os.path.join(GOLDEN, "small_model.npz"))


@pytest.fixture(scope="session")
def built_lib():
    """libmw_b200.so built in-tree (nvcc cross-compiles without a GPU)."""
    from manual_whisper_b200.build import build
    return build()


@pytest.fixture(scope="session")
def logmel_emu(tmp_path_factory):
    out = tmp_path_factory.mktemp("emu") / "logmel_emu"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", str(out), os.path.join(ROOT, "tests", "cpu_emu", "logmel_emu.cpp")])
    return str(out)
