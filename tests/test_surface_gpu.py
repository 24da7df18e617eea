"""Surface render mode on the GPU: the z-buffered splat (K9) bit-for-bit against the CPU oracle, the bilateral filter (K10)
and the lighting pass (K11) against the oracle and the reference's goldens, and the reference's own test scenarios
(tests/test_smooth.py, tests/test_render_output.py:448-556) through the drop-in classes."""
from pathlib import Path

import numpy as np
import numpy.testing as npt
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import c_oracle as co
from oracle import topsy_oracle as o
from test_surface_oracle import smooth_test_image

import topsy_b200 as topsy
from topsy_b200 import _native as N
from topsy_b200.canvas import offscreen
from topsy_b200.drawreason import DrawReason


@pytest.fixture(scope="module")
def surface_goldens():
    return np.load(Path(__file__).parent / "golden" / "surface_goldens.npz")


def _to_dev(*arrs):
    return [torch.from_numpy(np.ascontiguousarray(a, np.float32)).cuda() for a in arrs]


@pytest.mark.parametrize("R,scale,hmed", [(200, 30.0, 0.5), (256, 8.0, 0.05), (128, 2.0, 1.5), (333, 1.0, 0.02)],
                         ids=["mixed", "small", "huge-bilinear", "odd-res"])
def test_zbuffer_splat_bit_exact(R, scale, hmed):
    """Per pixel the fragment with the largest depth wins, whatever the order: the GPU image must equal the oracle's
    (max-depth rule) bit for bit -- inline footprints, deferred warp/CTA footprints and bilinear (>= 64 px) ones."""
    from topsy_b200.engine import SplatEngine
    rs = np.random.RandomState(11)
    n = 40000
    pos = rs.normal(size=(n, 3)).astype(np.float32) * np.float32(scale * 0.6)
    h = (hmed * np.exp(rs.normal(size=n) * 0.8)).astype(np.float32)
    m = rs.uniform(0.5, 1.5, n).astype(np.float32)
    q = rs.normal(size=n).astype(np.float32)
    rho = m / ((h * h) * h)
    cut = np.float32(np.quantile(rho, 0.3))
    M = o.transform_matrix(o.rotate(np.eye(3), 0.3, 0.4), np.array([0.1, -0.2, 0.05]), scale); sf = o.scale_factor(scale)
    lut = o.local_sphere_lut()
    ranges = (np.array([0, 10000, 25001], np.int64), np.array([9000, 15001, n - 25001], np.int64))
    want = co.splat_surface(pos[:, 0], pos[:, 1], pos[:, 2], h, m, q, M, sf, R, lut, cut, ranges=ranges, clamp_depth=False)
    assert (want[..., 1] > 0).mean() > 0.05

    eng = SplatEngine(R)
    try:
        eng.set_surface(lut, float(cut))
        eng.set_camera(M, sf)
        dev = _to_dev(pos[:, 0], pos[:, 1], pos[:, 2], h, m, q)
        eng.set_particles(*dev[:4]); eng.set_weights(dev[4], dev[5])
        # two calls = two progressive blocks: the second must keep the first one's z-buffer
        eng.render(N.MODE_SURFACE, ranges[0][:2], ranges[1][:2], clear=True)
        got = eng.render(N.MODE_SURFACE, ranges[0][2:], ranges[1][2:], clear=False).cpu().numpy()
    finally:
        eng.close()
    assert got.shape == (R, R, 2)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), \
        f"{(got != want).any(axis=2).sum()} pixels differ, max |d depth| {np.abs(got[..., 1] - want[..., 1]).max()}"


def test_surface_needs_its_lut():
    from topsy_b200.engine import SplatEngine
    eng = SplatEngine(64)
    try:
        dev = _to_dev(*[np.zeros(8)] * 6)
        eng.set_camera(np.eye(4), 1.0)
        eng.set_particles(*dev[:4]); eng.set_weights(dev[4], dev[5])
        with pytest.raises(RuntimeError):
            eng.render(N.MODE_SURFACE)
    finally:
        eng.close()


def test_smoothing_operation(surface_goldens):
    """The reference's tests/test_smooth.py, verbatim scenario and tolerances."""
    test_image = smooth_test_image()
    vis = topsy.test(100, render_resolution=test_image.shape[0], canvas_class=offscreen.VisualizerCanvas)
    vis.colormap.update_parameters({'type': 'surface', 'smoothing_scale': 0.02})
    surface_map = vis.colormap._impl
    smoothed_output = surface_map._smooth_numpy(test_image)
    npt.assert_allclose(test_image[..., 0], smoothed_output[..., 0], atol=1e-7)
    npt.assert_allclose(smoothed_output[::20, ::20, 1].ravel(), surface_goldens["test_smoothing_operation__expected_global_samples"],
                        atol=1e-6)
    npt.assert_allclose(smoothed_output[80:90, 80:90, 1].ravel(), surface_goldens["test_smoothing_operation__expected_edge_check"],
                        atol=1e-6)
    # and the whole image against the oracle
    want = o.bilateral_filter(test_image, *o.bilateral_params(0.02, test_image.shape[0]))
    npt.assert_allclose(smoothed_output, want, atol=2e-6)
    with pytest.raises(ValueError):
        surface_map._smooth_numpy(test_image[..., 0])


def test_surface_render(surface_goldens):
    """The reference's tests/test_render_output.py::test_surface_render, verbatim scenario and tolerances."""
    vis = topsy.test(int(1e5), render_resolution=200, canvas_class=offscreen.VisualizerCanvas, render_mode='surface')
    vis.quantity_name = "test-quantity"
    vis.scale = 30.0
    vis.rotate(0.0, 1.0)
    vis.render_sph(DrawReason.EXPORT)
    result = vis.get_sph_image()
    presentation_result = vis.get_sph_presentation_image()
    assert result.shape == (200, 200, 2)
    assert presentation_result.shape == (200, 200, 4)
    avoid_mask = np.ones(100, dtype=bool); avoid_mask[67] = False
    npt.assert_allclose(result[::20, ::20, 0].ravel()[avoid_mask],
                        surface_goldens["test_surface_render__quantity_expectation"][avoid_mask], rtol=1e-3)
    npt.assert_allclose(result[::20, ::20, 1].ravel(), surface_goldens["test_surface_render__depth_expectation"], rtol=1e-3)
    npt.assert_allclose(presentation_result[::20, ::20].ravel().astype(float),
                        surface_goldens["test_surface_render__presentation_expectation"], atol=30)
    # much tighter than the reference asks: the lit image agrees with the oracle-pinned golden to 2 counts
    assert np.abs(presentation_result[::20, ::20].ravel().astype(int)
                  - surface_goldens["test_surface_render__presentation_expectation"].astype(int)).max() <= 2
    # the colormap sits on a linear scale because the quantity is signed, and shows a colorbar (visualizer.py:327-328)
    assert vis.colormap.get_parameter('log') is False and vis.colormap.get_parameter('weighted_average') is True


def test_surface_density_cut_controls():
    vis = topsy.test(20000, render_resolution=128, canvas_class=offscreen.VisualizerCanvas, render_mode='surface')
    vis.scale = 30.0
    sph = vis._sph
    assert sph.get_density_cut_percentile_range() == (0.0, 100.0) and sph.get_density_cut_percentile() == 50.0
    vis.render_sph(DrawReason.EXPORT)
    covered_median = (sph.get_image()[..., 1] > 0).sum()
    sph.set_density_cut_percentile(0.0)
    vis.invalidate()
    vis.render_sph(DrawReason.EXPORT)
    covered_all = (sph.get_image()[..., 1] > 0).sum()
    assert covered_all > covered_median > 0
    tp = sph.last_transform_params
    assert np.float32(tp["density_cut"][0]) == np.float32(sph._percentile_to_den_cut[0])
    # no quantity selected: the material channel is identically zero and the surface is drawn without a colormap
    assert not sph.get_image()[..., 0].any()
    rgba = vis.get_sph_presentation_image()
    assert rgba.dtype == np.uint8 and rgba[..., 3].min() == 255 and rgba[..., :3].max() > 0


def test_surface_shade_matches_oracle_any_output_size():
    from topsy_b200.engine import SplatEngine
    rs = np.random.RandomState(5)
    R = 96
    yy, xx = np.mgrid[0:R, 0:R] / R
    img = np.zeros((R, R, 2), np.float32)
    img[..., 0] = np.exp(rs.normal(size=(R, R))).astype(np.float32)
    img[..., 1] = (0.3 + 0.2 * np.sin(6 * xx) * np.cos(5 * yy)).astype(np.float32)
    img[10:20, 30:50] = 0.0
    from topsy_b200.colormap import luts
    cm = luts.colormap_table_1d("viridis", 1000)
    eng = SplatEngine(R)
    try:
        for (ow, oh), colormap, log in [((R, R), True, True), ((150, 120), True, False), ((192, 256), False, False)]:
            p = N.SurfaceParams()
            p.depth_scale = 1.3
            p.light_direction[:] = [0.3, 0.5, 0.8]; p.light_color[:] = [1.0, 0.9, 0.8]; p.ambient_color[:] = [0.05, 0.0, 0.2]
            p.vmin, p.vmax = (-1.0, 1.0) if log else (0.0, 3.0)
            p.window_aspect_ratio = ow / oh
            p.material_colormap = int(colormap); p.log_scale = int(log)
            out = torch.empty((oh, ow, 4), dtype=torch.float32, device="cuda")
            eng.surface_shade(torch.from_numpy(img).cuda(), p, torch.from_numpy(cm).cuda() if colormap else None, out,
                              N.FMT_RGBA32F)
            want = o.surface_shade(img, ow, oh, depth_scale=np.float32(1.3), light_direction=(0.3, 0.5, 0.8),
                                   light_color=(1.0, 0.9, 0.8), ambient_color=(0.05, 0.0, 0.2),
                                   material_lut=cm if colormap else None, log=log, vmin=p.vmin, vmax=p.vmax)
            npt.assert_allclose(out.cpu().numpy(), want, atol=2e-4, err_msg=f"{ow}x{oh} colormap={colormap}")
    finally:
        eng.close()


def test_surface_autorange_on_device_matches_host():
    vis = topsy.test(50000, render_resolution=200, canvas_class=offscreen.VisualizerCanvas, render_mode='surface')
    vis.scale = 40.0
    vis.quantity_name = "test-quantity"        # autoranges on the device (Visualizer.use_device_autorange)
    dev = {k: vis.colormap.get_parameter(k) for k in ("vmin", "vmax", "log")}
    vis.colormap.autorange(vis._sph.get_image())   # the reference's host path on the same image
    host = {k: vis.colormap.get_parameter(k) for k in ("vmin", "vmax", "log")}
    assert dev["log"] == host["log"] is False
    npt.assert_allclose([dev["vmin"], dev["vmax"]], [host["vmin"], host["vmax"]], rtol=1e-6, atol=1e-12)
