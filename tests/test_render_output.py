"""The drop-in surface on the GPU against the reference's golden vectors -- same scenarios, calls and tolerances as the
reference's tests/test_render_output.py, with ``topsy_b200`` standing in for ``topsy``."""
import numpy as np
import numpy.testing as npt
import pytest

pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import topsy_b200 as topsy
from topsy_b200.canvas import offscreen
from topsy_b200.drawreason import DrawReason


@pytest.fixture(params=[False, True], ids=["nocells", "cells"])
def vis(request):
    vis = topsy.test(1000, render_resolution=200, canvas_class=offscreen.VisualizerCanvas, with_cells=request.param)
    vis.scale = 200.0
    return vis


def test_render(vis, goldens):
    result = vis.get_sph_presentation_image()
    assert result.dtype == np.uint8 and result.shape == (200, 200, 4)
    npt.assert_allclose(result[::20, ::20].ravel().astype(float), goldens["test_render__reference_result"], atol=5)


def test_hdr_rgb_render(goldens):
    vis = topsy.test(1000, render_resolution=200, canvas_class=offscreen.VisualizerCanvas, render_mode='rgb-hdr')
    vis.scale = 20.0
    vis.colormap.update_parameters({"min_mag": 38.0, "max_mag": 40.0})
    result = vis.get_sph_presentation_image()[..., :3]
    assert result.dtype == np.float16
    npt.assert_allclose(result[::20, ::20].ravel().astype(float), goldens["test_hdr_rgb_render__result_ref"], atol=1e-2)


def test_particle_pos_smooth(vis, goldens):
    if hasattr(vis.data_loader, '_cell_layout'):
        return
    npt.assert_allclose(vis.data_loader.get_pos_smooth()[::100], goldens["test_particle_pos_smooth___inline"], rtol=1e-6)


def test_sph_weighted_output(vis, goldens):
    vis.quantity_name = "test-quantity"
    vis.scale = 20.0
    vis.rotate(0.0, 0.4)
    vis.render_sph(DrawReason.EXPORT)
    result = vis.get_sph_image()
    assert result.shape == (200, 200)
    npt.assert_allclose(result[::20, ::20].flatten(), goldens["test_sph_weighted_output__expect"], atol=1.5e-7)


def test_sph_output(vis, goldens):
    vis.render_sph(DrawReason.EXPORT)
    result = vis.get_sph_image()
    assert result.shape == (200, 200)
    test = result[::20, ::20].flatten()
    expect = goldens["test_sph_output__expect"]
    npt.assert_allclose(test, expect, rtol=5e-1)
    assert abs((test / expect).mean() - 1.0) < 0.0015
    assert (test / expect).std() < 0.015


def test_periodic_sph_output(goldens):
    vis2 = topsy.test(1000, render_resolution=200, canvas_class=offscreen.VisualizerCanvas, periodic_tiling=True)
    vis2.scale = 200.0
    vis2.render_sph(DrawReason.EXPORT)
    result = vis2.get_sph_image()
    npt.assert_allclose(result[::20, ::20].flatten(), goldens["test_periodic_sph_output__expect"], rtol=1e-1)
    assert vis2.get_sph_presentation_image().shape == (200, 200, 4)


def test_periodic_accumulate_matches_oracle():
    """K7 against the numpy restatement of the replica sum, rotated view with fractional pixel shifts."""
    from oracle import topsy_oracle as o
    from topsy_b200 import periodic_sph
    vis2 = topsy.test(1000, render_resolution=200, canvas_class=offscreen.VisualizerCanvas, periodic_tiling=True)
    vis2.scale = 130.0
    vis2.rotate(0.3, 0.2)
    vis2.render_sph(DrawReason.EXPORT)
    base = vis2._sph._current_image().cpu().numpy().astype(np.float64)
    offs, wts = o.periodic_instances(vis2.rotation_matrix, 100.0 / 130.0)
    offs2, wts2 = periodic_sph.replica_offsets_and_weights(vis2.rotation_matrix, 100.0 / 130.0)
    assert len(wts) == len(wts2)
    order = np.lexsort(np.round(offs.T, 4)); order2 = np.lexsort(np.round(offs2.T, 4))
    npt.assert_allclose(offs[order], offs2[order2], atol=1e-6); npt.assert_allclose(wts[order], wts2[order2], atol=1e-6)
    want = o.periodic_accumulate(base, offs, wts)
    got = vis2._sph._current_periodic_image().cpu().numpy()
    big = want[..., 0] > 1e-6 * want[..., 0].max()
    npt.assert_allclose(got[..., 0][big], want[..., 0][big], rtol=2e-4)


def test_rotated_sph_output(vis):
    vis.draw(reason=DrawReason.EXPORT)
    unrotated = vis.get_sph_image()
    vis.rotation_matrix = np.array([[0.0, 1.0, 0.0], [-1.0, 0.0, 0.0], [0.0, 0.0, 1.0]], dtype=np.float32)
    vis.draw(reason=DrawReason.EXPORT)
    npt.assert_allclose(unrotated.T[:, ::-1], vis.get_sph_image(), rtol=5e-2)


def test_rgb_sph_output():
    vis = topsy.test(1000, render_resolution=200, canvas_class=offscreen.VisualizerCanvas, render_mode='rgb')
    assert vis.get_sph_image().shape == (200, 200, 3)
    assert vis._sph.get_image().shape == (200, 200, 4)


def test_depth_output(goldens):
    vis = topsy.test(1000, render_resolution=200, canvas_class=offscreen.VisualizerCanvas)
    vis.scale = 20.0
    vis.rotation_matrix = np.array([[1.0, 0.0, 0.0], [0.0, 0.0, 1.0], [0.0, -1.0, 0.0]], dtype=np.float32)
    vis.render_sph(DrawReason.EXPORT)
    result = vis._sph.get_depth_image(DrawReason.EXPORT)
    npt.assert_allclose(result[::20, ::20].ravel(), goldens["test_depth_output__expect"], atol=1e-1)


def test_bivariate_render(goldens):
    vis = topsy.test(1000, render_resolution=200, canvas_class=offscreen.VisualizerCanvas, render_mode='bivariate')
    vis.quantity_name = "test-quantity"
    vis.scale = 20.0
    vis.rotate(0.0, 0.5)
    vis.render_sph(DrawReason.EXPORT)
    results = vis.get_sph_image()
    mapped = vis.get_sph_presentation_image()
    npt.assert_allclose(results[::20, ::20, 0].ravel(), goldens["test_bivariate_render__expect_den"], rtol=2e-3)
    npt.assert_allclose(results[::20, ::20, 1].ravel(), goldens["test_bivariate_render__expect_qty"], atol=1e-4)
    got = mapped[::20, ::20].ravel().astype(float)
    # matplotlib's 510-entry twilight_shifted vs OpenCV's 256-entry copy: a few LUT cells differ by more than 5
    assert np.mean(np.abs(got - goldens["test_bivariate_render__expect_rgba"]) <= 5) > 0.97


def test_presentation_image_and_canvas_draw(vis):
    img = vis.get_presentation_image((320, 240))
    assert img.shape == (240, 320, 4) and img.dtype == np.uint8
    frame = vis.canvas.draw()
    assert frame.shape == (480, 640, 4)
    assert frame[..., 3].min() == 255


def test_progressive_refine_converges_to_export():
    """Interactive CHANGE frame + REFINE frames accumulate to the EXPORT image (progression + cells + mass scale)."""
    vis = topsy.test(200000, render_resolution=256, canvas_class=offscreen.VisualizerCanvas, with_cells=True)
    vis.scale = 40.0
    vis.render_sph(DrawReason.EXPORT)
    full = vis._sph.get_image()[..., 0].copy()
    vis._sph._render_progression._recommended_num_particles_to_render = 20000

    class SlowClock:
        """Pretends every block takes a whole frame budget, as on a slow GPU (a B200 finishes 200k particles at once)."""
        def __init__(self):
            self.t = 0.0
            self.running_mean_duration = 1.0 / 30
        def __enter__(self):
            return self
        def __exit__(self, *exc):
            self.t += 1.0 / 30
        def total_time_in_frame(self, wait=True):
            return self.t
        def end_frame(self):
            self.t = 0.0

    vis._sph._render_timer = SlowClock()
    vis.render_sph(DrawReason.CHANGE)
    assert vis._sph.needs_refine()
    first_scale = vis._sph.last_render_mass_scale
    assert first_scale > 1.5
    coarse = vis._sph.get_image()[..., 0]
    big = full > 1e-3 * full.max()
    assert abs(np.median(coarse[big] / full[big]) - 1.0) < 0.2           # a fair subsample, rescaled
    frames = 0
    while vis._sph.needs_refine() and frames < 200:
        vis.render_sph(DrawReason.REFINE)
        frames += 1
    assert not vis._sph.needs_refine()
    assert vis._sph.last_render_mass_scale == pytest.approx(1.0)
    npt.assert_allclose(vis._sph.get_image()[..., 0][big], full[big], rtol=1e-4)
