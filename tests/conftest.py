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


def pytest_collection_modifyitems(config, items):
    """The oracle-vs-golden pins (test_oracle_golden.py) are CPU tests; on a CUDA box they are
    ALSO given the `gpu` marker, so that the driver's `-m gpu` run re-checks, on the machine that
    produces the parity numbers, that the checker itself still matches the reference's outputs.
    Without a GPU they stay unmarked and run in the `-m "not gpu"` suite."""
    try:
        import torch
        has_cuda = torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        has_cuda = False
    if not has_cuda:
        return
    for item in items:
        if "test_oracle_golden" in item.nodeid:
            item.add_marker(pytest.mark.gpu)


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
