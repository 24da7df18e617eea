import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """gpu-marked tests are skipped (not failed) where there is no CUDA device or the native library was not built."""
    reason = None
    try:
        import torch
        if not torch.cuda.is_available():
            reason = "needs a CUDA device"
    except Exception:
        reason = "needs torch with CUDA"
    if reason is None and not (ROOT / "topsy_b200" / "libtsplat.so").exists():
        reason = "topsy_b200/libtsplat.so not built (python -c 'import __graft_entry__ as g; g.build()')"
    if reason is None:
        return
    skip = pytest.mark.skip(reason=reason)
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


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
