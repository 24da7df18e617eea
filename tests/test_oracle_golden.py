"""Pins the CPU oracle (oracle/topsy_oracle.py, oracle/splat_oracle.c) to the reference's own golden vectors
(tests/test_render_output.py of the reference, extracted into tests/golden/reference_goldens.npz) with the reference's
own tolerances, and the numpy and C restatements to each other.  No GPU needed."""
import numpy as np
import pytest

from oracle import c_oracle as co
from oracle import topsy_oracle as o

R = 200


@pytest.fixture(scope="module")
def fx():
    return o.GMMFixture(1000)


def render(fx, lut, scale, rot, mode, weights, impl=co):
    ps = fx.pos_smooth()
    M = o.transform_matrix(rot, np.zeros(3), scale)
    return impl.splat(ps[:, 0], ps[:, 1], ps[:, 2], ps[:, 3], weights, M, o.scale_factor(scale), R, mode, lut)


def test_fixture_data_known_answer(fx, goldens, sched_goldens):
    # reference tests/test_render_output.py:144-159
    np.testing.assert_allclose(fx.pos_smooth()[::100], goldens["test_particle_pos_smooth___inline"], rtol=1e-6)
    # and bit-for-bit against the reference's own TestDataLoader run in the build container
    assert np.array_equal(fx.pos_smooth(), sched_goldens["gmm1000_pos_smooth"])
    assert np.array_equal(fx.mass, sched_goldens["gmm1000_mass"])
    assert np.array_equal(fx.quantity, sched_goldens["gmm1000_qty"])
    assert np.array_equal(fx.rgb, sched_goldens["gmm1000_rgb"])
    assert np.array_equal(o.GMMFixture(1).pos_smooth(), sched_goldens["gmm1_pos_smooth"])


def test_density_golden(fx, goldens, oracle_lut):
    # reference :200-241 (vis.scale = 200, density)
    img = render(fx, oracle_lut, 200.0, np.eye(3), o.MODE_WEIGHTED, (fx.mass, np.zeros(1000, np.float32)))
    test = img[::20, ::20, 0].ravel()
    expect = goldens["test_sph_output__expect"]
    np.testing.assert_allclose(test, expect, rtol=5e-1)
    assert abs((test / expect).mean() - 1.0) < 0.0015
    assert (test / expect).std() < 0.015


def test_weighted_golden(fx, goldens, oracle_lut):
    # reference :161-198 (scale 20, rotate(0, 0.4), test-quantity)
    img = render(fx, oracle_lut, 20.0, o.rotate(np.eye(3), 0.0, 0.4), o.MODE_WEIGHTED, (fx.mass, fx.quantity.astype(np.float32)))
    res = (img[..., 1] / img[..., 0])[::20, ::20].ravel()
    np.testing.assert_allclose(res, goldens["test_sph_weighted_output__expect"], atol=1.5e-7)


def test_bivariate_golden(fx, goldens, oracle_lut):
    # reference :345-446
    img = render(fx, oracle_lut, 20.0, o.rotate(np.eye(3), 0.0, 0.5), o.MODE_WEIGHTED, (fx.mass, fx.quantity.astype(np.float32)))
    np.testing.assert_allclose(img[::20, ::20, 0].ravel(), goldens["test_bivariate_render__expect_den"], rtol=2e-3)
    np.testing.assert_allclose((img[..., 1] / img[..., 0])[::20, ::20].ravel(), goldens["test_bivariate_render__expect_qty"], atol=1e-4)


def test_rgb_hdr_golden(fx, goldens, oracle_lut):
    # reference :69-141 (rgb-hdr, scale 20, min_mag 38, max_mag 40) -- includes the tri-band colormap stage
    rgb = fx.rgb
    img = render(fx, oracle_lut, 20.0, np.eye(3), o.MODE_RGB, (rgb[:, 0], rgb[:, 1], rgb[:, 2]))
    p = o.colormap_params(o.mag_per_arcsec2_to_log_output(40.0), o.mag_per_arcsec2_to_log_output(38.0), True, False, 1.0,
                          may_produce_weighted_average=False)
    out = o.colormap_rgb(img, p)[..., :3].astype(np.float16)
    np.testing.assert_allclose(out[::20, ::20].ravel().astype(np.float64), goldens["test_hdr_rgb_render__result_ref"], atol=1e-2)


def test_magnitude_conversion_known_answers():
    # reference tests/test_colormap.py:88-104: vmin/vmax 1, 2 <-> max_mag/min_mag 34.07.., 31.57..; mags 1, 2 <-> 14.22.., 13.82..
    assert np.isclose(o.mag_per_arcsec2_to_log_output(31.57212566586528), 2.0)
    assert np.isclose(o.mag_per_arcsec2_to_log_output(34.07212566586528), 1.0)
    assert np.isclose(o.mag_per_arcsec2_to_log_output(2.0), 13.828850266346112)
    assert np.isclose(o.mag_per_arcsec2_to_log_output(1.0), 14.228850266346113)


def test_depth_golden(fx, goldens, oracle_lut):
    # reference :302-343
    rot = np.array([[1.0, 0, 0], [0, 0, 1.0], [0, -1.0, 0]])
    img = render(fx, oracle_lut, 20.0, rot, o.MODE_DEPTH, (fx.mass,))
    depth = (img[..., 1] / img[..., 0] - 0.5) * 20.0 * 2.0
    np.testing.assert_allclose(depth[::20, ::20].ravel(), goldens["test_depth_output__expect"], atol=1e-1)


def test_rotation_identity(fx, oracle_lut):
    # reference :280-293
    w = (fx.mass, np.zeros(1000, np.float32))
    a = render(fx, oracle_lut, 200.0, np.eye(3), o.MODE_WEIGHTED, w)[..., 0]
    b = render(fx, oracle_lut, 200.0, np.array([[0, 1.0, 0], [-1.0, 0, 0], [0, 0, 1.0]]), o.MODE_WEIGHTED, w)[..., 0]
    np.testing.assert_allclose(a.T[:, ::-1], b, rtol=5e-2)


def test_density_rgba_golden_with_twilight_shifted(fx, goldens, oracle_lut):
    """reference :27-65 (test_render): density + autorange + log + twilight_shifted LUT, RGBA8 atol 5.  The LUT values
    come from OpenCV's copy of matplotlib's table when matplotlib is absent (topsy_b200/colormap/luts.py)."""
    from topsy_b200.colormap import luts
    img = render(fx, oracle_lut, 200.0, np.eye(3), o.MODE_WEIGHTED, (fx.mass, np.zeros(1000, np.float32)))
    auto = o.autorange_scalar(img[..., 0].astype(np.float32))
    assert auto["log"]
    p = o.colormap_params(auto["vmin"], auto["vmax"], True, False, 1.0)
    rgba = o.to_unorm8(o.colormap_scalar(img, p, luts.colormap_table_1d("twilight_shifted", 1000), True, False))
    np.testing.assert_allclose(rgba[::20, ::20].ravel().astype(float), goldens["test_render__reference_result"], atol=5)


def test_bivariate_rgba_golden(fx, goldens, oracle_lut):
    """reference :345-359,412-446: bivariate 2-D LUT presentation, RGBA8 atol 5.  The reference test sets the quantity
    (which autoranges at the *initial* view: scale 200, no rotation) before zooming to scale 20, so the colormap ranges
    are the stale wide-view ones -- reproduced here."""
    from topsy_b200.colormap import luts
    w = (fx.mass, fx.quantity.astype(np.float32))
    wide = render(fx, oracle_lut, 200.0, np.eye(3), o.MODE_WEIGHTED, w)
    with np.errstate(divide="ignore", invalid="ignore"):
        dlog = np.log10(wide[..., 0].astype(np.float32)); dlog = dlog[np.isfinite(dlog)]
        dvmin, dvmax = np.percentile(dlog, [1.0, 99.9])
        auto = o.autorange_scalar((wide[..., 1].astype(np.float32) / wide[..., 0].astype(np.float32)))
    assert not auto["log"]          # the test quantity is signed
    img = render(fx, oracle_lut, 20.0, o.rotate(np.eye(3), 0.0, 0.5), o.MODE_WEIGHTED, w)
    p = o.colormap_params(auto["vmin"], auto["vmax"], False, True, 1.0, density_vmin=dvmin, density_vmax=dvmax)
    rgba = o.to_unorm8(o.colormap_bivariate(img, p, luts.colormap_table_2d("twilight_shifted", 1000), False, True))
    got = rgba[::20, ::20].ravel().astype(float)
    want = goldens["test_bivariate_render__expect_rgba"]
    # the reference's LUT is matplotlib's 510-entry table, ours may be OpenCV's 256-entry copy: allow a few outliers
    assert np.mean(np.abs(got - want) <= 5) > 0.97, np.abs(got - want).max()


@pytest.mark.parametrize("mode,scale,angles", [(o.MODE_DENSITY, 200.0, (0, 0)), (o.MODE_WEIGHTED, 20.0, (0, 0.5)),
                                               (o.MODE_RGB, 5.0, (0.3, 0.4)), (o.MODE_DEPTH, 1.0, (1.0, -0.7)),
                                               (o.MODE_WEIGHTED, 0.3, (0.1, 0.2))])
def test_numpy_and_c_restatements_agree(fx, oracle_lut, mode, scale, angles):
    rgb = fx.rgb
    w = {o.MODE_DENSITY: (fx.mass,), o.MODE_DEPTH: (fx.mass,), o.MODE_WEIGHTED: (fx.mass, fx.quantity.astype(np.float32)),
         o.MODE_RGB: (rgb[:, 0], rgb[:, 1], rgb[:, 2])}[mode]
    rot = o.rotate(np.eye(3), *angles)
    a = render(fx, oracle_lut, scale, rot, mode, w, impl=o)
    b = render(fx, oracle_lut, scale, rot, mode, w, impl=co)
    scale_ = np.abs(a).max()
    assert np.abs(a - b).max() <= 1e-12 * scale_


def test_ceil_bounds_equal_comparisons():
    """The CUDA kernels turn the coverage comparisons (j + 0.5 >= e0) & (j + 0.5 < e1) into integer ranges with ceil();
    check the two forms agree for every pixel of the image, including fp32 edge cases around *.5 and huge values."""
    rs = np.random.RandomState(3)
    Rr = 64
    e = np.concatenate([rs.uniform(-3, Rr + 3, 20000), np.arange(-2, Rr + 2) + 0.5, np.arange(-2, Rr + 2) + 0.5 + 1e-6,
                        np.arange(-2, Rr + 2) + 0.5 - 1e-6, [-1e30, 1e30, -0.49999997, 0.49999997, 1e-10]]).astype(np.float32)
    e1 = (e + rs.uniform(0, 9, len(e)).astype(np.float32)).astype(np.float32)
    j = np.arange(Rr, dtype=np.float32) + np.float32(0.5)
    cover = (j[None, :] >= e[:, None]) & (j[None, :] < e1[:, None])
    lo = np.maximum(np.ceil(e - np.float32(0.5)), 0)
    hi = np.minimum(np.ceil(e1 - np.float32(0.5)) - 1, Rr - 1)
    jj = np.arange(Rr, dtype=np.float32)
    cover2 = (jj[None, :] >= lo[:, None]) & (jj[None, :] <= hi[:, None])
    assert np.array_equal(cover, cover2)


def test_kernel_lut_is_mass_conserving_and_product_lut_matches(oracle_lut):
    from topsy_b200.kernel_lut import kernel_lut
    for n, off in zip(o.LUT_LEVEL_SIZES, o.LUT_LEVEL_OFFSETS):
        level = oracle_lut[off:off + n * n].astype(np.float64)
        assert abs(level.sum() * (4.0 / n) ** 2 - 1.0) < 1e-6
    prod = kernel_lut()
    assert prod.shape == (5440,) and prod.dtype == np.float32
    np.testing.assert_allclose(prod, oracle_lut, rtol=2e-7, atol=1e-12)
