import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def goldens():
    import numpy as np
    return np.load(ROOT / "tests" / "golden" / "reference_goldens.npz")


@pytest.fixture(scope="session")
def sched_goldens():
    import numpy as np
    return np.load(ROOT / "tests" / "golden" / "scheduling_goldens.npz")


@pytest.fixture(scope="session")
def oracle_lut():
    from oracle import topsy_oracle as o
    return o.kernel_lut()
