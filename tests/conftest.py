import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    path = os.path.join(ROOT, "tests", "golden", "golden_v1.npz")
    data = np.load(path, allow_pickle=False)
    return data


@pytest.fixture(scope="session")
def golden_backdoor():
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_backdoor_v1.npz"), allow_pickle=False)


def golden_names(kind=None):
    path = os.path.join(ROOT, "tests", "golden", "golden_v1.npz")
    data = np.load(path, allow_pickle=False)
    names = [str(x) for x in data["__names__"]]
    if kind is not None:
        names = [n for n in names if f"{n}/{kind}" in data.files]
    return names
