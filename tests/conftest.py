import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")

# FZ_EMULATED_DEVICE=1: the `-m gpu` tests run with the library compiled for the CUDA-on-CPU emulator behind the product's Python layers
# (tests/emu_device.py); used by tests/test_gpu_suite_on_emulator.py, which runs the lighter GPU test files that way in the CPU suite
if os.environ.get("FZ_EMULATED_DEVICE") == "1":
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import emu_device

    emu_device.install(emu_device.Permanent)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_json():
    with open(os.path.join(GOLDEN, "golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def ntt_golden():
    return dict(np.load(os.path.join(GOLDEN, "ntt_golden.npz")))


@pytest.fixture(scope="session")
def scheme_golden():
    return dict(np.load(os.path.join(GOLDEN, "scheme_golden.npz")))


@pytest.fixture(scope="session")
def kat_golden():
    return dict(np.load(os.path.join(GOLDEN, "kat_golden.npz")))


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as orc

    orc.lib()
    return orc


@pytest.fixture(scope="session")
def ring_golden():
    return dict(np.load(os.path.join(GOLDEN, "ring_golden.npz")))
