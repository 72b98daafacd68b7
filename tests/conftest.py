import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_golden.npz"))


@pytest.fixture(scope="session")
def seed_golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_seed_golden.npz"))


@pytest.fixture(scope="session")
def lib_built():
    from wann_b200 import build
    return build.build()
