"""Pins the surface-mode oracle (oracle_splat_surface in oracle/splat_oracle.c; bilateral filter and lighting in
oracle/topsy_oracle.py) to the reference's golden vectors: tests/test_smooth.py::test_smoothing_operation (atol 1e-6) and
tests/test_render_output.py::test_surface_render (:448-556; rtol 1e-3 on the smoothed (quantity, depth) image, atol 30 on
the lit RGBA), extracted into tests/golden/surface_goldens.npz by tests/golden/make_golden.py.  No GPU needed."""
import numpy as np
import pytest

from oracle import c_oracle as co
from oracle import topsy_oracle as o

R = 200


@pytest.fixture(scope="module")
def surface_goldens():
    from pathlib import Path
    return np.load(Path(__file__).parent / "golden" / "surface_goldens.npz")


def smooth_test_image(width=256, height=256):
    """The input of the reference's tests/test_smooth.py:11-47 (seeded legacy numpy stream)."""
    np.random.seed(1337)
    X, Y = np.meshgrid(np.linspace(0, 1, width), np.linspace(0, 1, height))
    img = np.zeros((height, width, 2), dtype=np.float32)
    step = np.zeros_like(X); step[height // 4:3 * height // 4, width // 4:3 * width // 4] = 0.5
    img[:, :, 0] = X * 0.5 + Y * 0.3 + step + np.random.normal(0, 0.05, (height, width))
    step2 = np.zeros_like(X); step2[height // 3:2 * height // 3, width // 3:2 * width // 3] = 0.3
    img[:, :, 1] = Y * 0.4 + X * 0.2 + step2 + np.random.normal(0, 0.03, (height, width))
    return np.abs(img) + 0.01


def test_bilateral_filter_golden(surface_goldens):
    img = smooth_test_image()
    spatial, rng, ksize = o.bilateral_params(0.02, img.shape[0])
    assert ksize == 21
    out = o.bilateral_filter(img, spatial, rng, ksize)
    np.testing.assert_allclose(out[..., 0], img[..., 0], atol=1e-7)
    np.testing.assert_allclose(out[::20, ::20, 1].ravel(), surface_goldens["test_smoothing_operation__expected_global_samples"],
                               atol=1e-6)
    np.testing.assert_allclose(out[80:90, 80:90, 1].ravel(), surface_goldens["test_smoothing_operation__expected_edge_check"],
                               atol=1e-6)


@pytest.fixture(scope="module")
def scene():
    fx = o.GMMFixture(100000)
    cut = np.float32(o.density_cut_value(o.density_cut_table(fx.mass, fx.smooth), 50.0))
    return fx, fx.pos_smooth(), fx.quantity.astype(np.float32), cut, o.local_sphere_lut()


def surface_render(scene, scale, rot, clamp_depth):
    fx, ps, q, cut, lut = scene
    M = o.transform_matrix(rot, np.zeros(3), scale)
    return co.splat_surface(ps[:, 0], ps[:, 1], ps[:, 2], ps[:, 3], fx.mass, q, M, o.scale_factor(scale), R, lut, cut,
                            clamp_depth=clamp_depth)


@pytest.mark.parametrize("clamp_depth", [True, False], ids=["reference-depth-test", "max-depth"])
def test_surface_render_golden(scene, surface_goldens, clamp_depth):
    img = surface_render(scene, 30.0, o.rotate(np.eye(3), 0.0, 1.0), clamp_depth)
    out = o.bilateral_filter(img, *o.bilateral_params(0.01, R))
    keep = np.ones(100, bool); keep[67] = False         # the reference excludes this pixel too (:512-515)
    np.testing.assert_allclose(out[::20, ::20, 0].ravel()[keep], surface_goldens["test_surface_render__quantity_expectation"][keep],
                               rtol=1e-3)
    np.testing.assert_allclose(out[::20, ::20, 1].ravel(), surface_goldens["test_surface_render__depth_expectation"], rtol=1e-3)


def test_surface_presentation_golden(scene, surface_goldens):
    from topsy_b200 import config
    from topsy_b200.colormap import luts
    # visualizer.py:299-338: the material range is autoranged when the quantity is selected, i.e. on the default view
    first = surface_render(scene, config.DEFAULT_SCALE, np.eye(3), True)
    vals = first[..., 0].ravel()[first[..., 1].ravel() > 0.0]           # surface.py:256-259
    assert (vals < 0).any()                                              # -> linear scale
    vmin, vmax = np.percentile(vals, [1.0, 99.9])
    img = surface_render(scene, 30.0, o.rotate(np.eye(3), 0.0, 1.0), True)
    out = o.bilateral_filter(img, *o.bilateral_params(0.01, R))
    rgba = o.to_unorm8(o.surface_shade(out, R, R, material_lut=luts.colormap_table_1d(config.DEFAULT_COLORMAP, 1000), log=False,
                                       vmin=np.float32(vmin), vmax=np.float32(vmax)))
    np.testing.assert_allclose(rgba[::20, ::20].ravel().astype(float), surface_goldens["test_surface_render__presentation_expectation"],
                               atol=30)
    assert np.abs(rgba[::20, ::20].ravel().astype(int) - surface_goldens["test_surface_render__presentation_expectation"].astype(int)).max() <= 2


def test_local_sphere_lut_shape_and_values():
    lut = o.local_sphere_lut()
    assert lut.shape == (o.LUT_TOTAL,) and lut.dtype == np.float32
    lvl3 = lut[o.LUT_LEVEL_OFFSETS[3]:].reshape(8, 8)
    assert lvl3[0, 0] == np.float32(-0.01) and lvl3[3, 3] == np.float32(np.sqrt(4 - 2 * 0.25 ** 2))


@pytest.mark.parametrize("clamp_depth", [True, False])
@pytest.mark.parametrize("scale", [30.0, 4.0])
def test_numpy_and_c_surface_restatements_agree(clamp_depth, scale):
    """Bit for bit, including a zoomed view where depths exceed 1.0 (where the two depth rules differ from each other)."""
    fx = o.GMMFixture(1500)
    ps = fx.pos_smooth()
    q = fx.quantity.astype(np.float32)
    cut = np.float32(o.density_cut_value(o.density_cut_table(fx.mass, fx.smooth), 30.0))
    lut = o.local_sphere_lut()
    M = o.transform_matrix(o.rotate(np.eye(3), 0.2, 0.7), np.zeros(3), scale); sf = o.scale_factor(scale)
    a = o.splat_surface(ps[:, 0], ps[:, 1], ps[:, 2], ps[:, 3], fx.mass, q, M, sf, 96, lut, cut, clamp_depth=clamp_depth)
    b = co.splat_surface(ps[:, 0], ps[:, 1], ps[:, 2], ps[:, 3], fx.mass, q, M, sf, 96, lut, cut, clamp_depth=clamp_depth)
    assert (b[..., 1] > 0).mean() > 0.1
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_depth_rules_differ_only_beyond_unit_depth():
    fx = o.GMMFixture(1500)
    ps = fx.pos_smooth(); q = fx.quantity.astype(np.float32)
    lut = o.local_sphere_lut()
    M = o.transform_matrix(np.eye(3), np.zeros(3), 4.0); sf = o.scale_factor(4.0)
    ref = co.splat_surface(ps[:, 0], ps[:, 1], ps[:, 2], ps[:, 3], fx.mass, q, M, sf, 96, lut, 0.0, clamp_depth=True)
    mx = co.splat_surface(ps[:, 0], ps[:, 1], ps[:, 2], ps[:, 3], fx.mass, q, M, sf, 96, lut, 0.0, clamp_depth=False)
    differ = (ref != mx).any(axis=2)
    assert differ.any()                                   # the zoomed view does push fragments beyond depth 1
    assert (mx[..., 1][differ] > 1.0).all()               # ... and only those pixels differ
    assert (mx[..., 1] >= ref[..., 1]).all()
