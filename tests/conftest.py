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


def load_golden(name):
    """{case: {field: ndarray}} from tests/golden/<name>.npz."""
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    cases = {}
    for key in z.files:
        case, field = key.split("/", 1)
        cases.setdefault(case, {})[field] = z[key]
    return cases


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
