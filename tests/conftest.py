import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """Build (or reuse) the in-tree CUDA library and the test-only CPU twin."""
    import __graft_entry__ as g
    return g.build()


@pytest.fixture(scope="session")
def golden_v2():
    import numpy as np
    return np.load(os.path.join(GOLDEN, "sbr_v2_cases.npz"))


@pytest.fixture(scope="session")
def golden_v2_tight():
    import numpy as np
    return np.load(os.path.join(GOLDEN, "sbr_v2_tight.npz"))


@pytest.fixture(scope="session")
def stage_samples():
    import numpy as np
    return np.load(os.path.join(GOLDEN, "stage_samples.npz"))


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("this test is marked gpu but no CUDA device is visible (there is no CPU fallback)")
    return torch.device("cuda:0")
