import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


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
