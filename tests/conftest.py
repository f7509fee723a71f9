import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs the reference tree at /root/reference (build container)")


@pytest.fixture(scope="session")
def golden_demo():
    import numpy as np
    return dict(np.load(os.path.join(GOLDEN, "demo_cfg1.npz")))


@pytest.fixture(scope="session")
def golden_random():
    import numpy as np
    return dict(np.load(os.path.join(GOLDEN, "random_beta.npz")))


@pytest.fixture(scope="session")
def golden_extras():
    import numpy as np
    return dict(np.load(os.path.join(GOLDEN, "extras.npz")))
