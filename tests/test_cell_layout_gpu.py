"""K4: device CellLayout.from_positions against the reference's own numpy module (golden) and the host path."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from topsy_b200 import config
from topsy_b200.cell_layout import CellLayout


@pytest.mark.parametrize("tag", ["cl64", "cl32"])
def test_device_layout_is_bit_exact(sched_goldens, tag):
    g = sched_goldens
    pos = g[f"{tag}_pos"]
    if tag == "cl64":
        bmin, bmax, nside = -1.0, 1.0, 10
    else:
        bmin, bmax = g["cl32_box"].astype(np.float32)
        nside = config.DEFAULT_CELLS_NSIDE
    layout, order = CellLayout.from_positions(torch.from_numpy(pos).cuda(), bmin, bmax, nside)
    assert np.array_equal(layout._lengths, g[f"{tag}_lengths"])
    assert np.array_equal(layout._offsets, g[f"{tag}_offsets"])
    assert np.array_equal(layout._centres, g[f"{tag}_centres"])
    host_layout, host_order = CellLayout.from_positions(pos, bmin, bmax, nside)
    assert np.array_equal(order.cpu().numpy(), host_order)                  # both are the stable argsort


def test_large_random_and_errors():
    rs = np.random.RandomState(9)
    pos = rs.uniform(-50, 50, (3_000_017, 3)).astype(np.float32)
    lo = pos.min(); hi = pos.max(); pad = np.float32(1e-5) * (hi - lo)
    layout, order = CellLayout.from_positions(torch.from_numpy(pos).cuda(), lo - pad, hi + pad, 16)
    host_layout, host_order = CellLayout.from_positions(pos, lo - pad, hi + pad, 16)
    assert np.array_equal(layout._lengths, host_layout._lengths)
    assert np.array_equal(order.cpu().numpy(), host_order)
    with pytest.raises(ValueError):
        CellLayout.from_positions(torch.from_numpy(pos).cuda(), -10.0, 10.0, 16)
    empty_layout, empty_order = CellLayout.from_positions(torch.zeros((1, 3), device="cuda"), -1.0, 1.0, 4)
    assert empty_layout._lengths.sum() == 1 and empty_order.cpu().numpy().tolist() == [0]
