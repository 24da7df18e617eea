"""Periodic tiling host logic + oracle pin (no GPU): the numpy restatement of PeriodicSPH reproduces the reference's
golden image (tests/test_render_output.py:243-278 of the reference, rtol 1e-1)."""
import numpy as np

from oracle import c_oracle as co
from oracle import topsy_oracle as o


def test_periodic_oracle_matches_reference_golden(goldens, oracle_lut):
    fx = o.GMMFixture(1000)
    ps = fx.pos_smooth()
    M = o.transform_matrix(np.eye(3), np.zeros(3), 200.0)
    img = co.splat(ps[:, 0], ps[:, 1], ps[:, 2], ps[:, 3], (fx.mass, np.zeros(1000, np.float32)), M, o.scale_factor(200.0), 200,
                   o.MODE_WEIGHTED, oracle_lut)
    offs, w = o.periodic_instances(np.eye(3), 100.0 / 200.0)
    assert len(w) == 25                      # 5 x 5 in-plane replicas of the single depth layer |z| < 1
    out = o.periodic_accumulate(img, offs, w)
    np.testing.assert_allclose(out[::20, ::20, 0].ravel(), goldens["test_periodic_sph_output__expect"], rtol=1e-1)


def test_replica_weights_fade_with_depth():
    rot = o.rotate(np.eye(3), 0.0, 0.6)
    offs, w = o.periodic_instances(rot, 1.0)
    assert (w > 0).all() and (w <= 1).all() and (w < 1).any()
