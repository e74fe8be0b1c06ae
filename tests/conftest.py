import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    gdir = os.path.join(ROOT, "tests", "golden")
    return {name: np.load(os.path.join(gdir, name + ".npz")) for name in ("hrnet_forward", "lanczos", "cpsnr", "predict", "loss", "apply_shifts", "evaluate", "shiftnet", "trainstep")}
