import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "pending_gpu: GPU test written after a round's GPU budget was spent -- not yet run on a GPU; "
                                       "run with -m pending_gpu (tools/gpu_round.sh does) and promote to `gpu` once green")


@pytest.fixture(scope="session")
def golden():
    """Outputs of the real reference (tests/golden/make_golden.py, run in the build container)."""
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_golden.npz"))
