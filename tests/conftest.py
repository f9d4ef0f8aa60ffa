"""Shared fixtures. `-m "not gpu"`: oracle, host logic, ABI surface. `-m gpu`: CUDA-vs-oracle parity through the C ABI."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


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
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    pyoracle.lib()
    return pyoracle


@pytest.fixture(scope="session")
def stream4():
    """Four consecutive 640x480 frames of the cfg-2 (fr1/xyz-shape) synthetic stream."""
    from lineslam_b200 import synth
    imgs, deps, poses = synth.make_stream(4, scene_seed=2000)
    return imgs, deps, poses, synth.camera_K()


@pytest.fixture(scope="session")
def small_frames():
    """Two 320x240 frames (ragged sizes are covered separately)."""
    from lineslam_b200 import synth
    imgs, deps, poses = synth.make_stream(2, scene_seed=2001, W=320, H=240)
    return imgs, deps, poses, synth.camera_K(320, 240)


@pytest.fixture(scope="session")
def api():
    from lineslam_b200 import api as a
    a.lib()
    return a
