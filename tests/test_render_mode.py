"""Render-mode switching on the GPU (reference: tests/test_render_mode.py, same scenarios)."""
import numpy as np
import pytest

pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import topsy_b200 as topsy
from topsy_b200.canvas import offscreen

MODES = ['univariate', 'bivariate', 'rgb', 'rgb-hdr', 'surface']


def check_outputs(vis, mode):
    result = vis.get_sph_image()
    pres = vis.get_sph_presentation_image()
    assert pres.dtype == (np.float16 if mode.endswith('hdr') else np.uint8)
    res = vis._render_resolution
    assert pres.shape == (res, res, 4)
    if mode in ('rgb', 'rgb-hdr'):
        assert result.shape == (res, res, 3)
    elif mode in ('bivariate', 'surface'):
        assert result.shape == (res, res, 2)
    else:
        assert result.shape == (res, res)


def test_switching():
    vis = topsy.test(1000, render_resolution=200, canvas_class=offscreen.VisualizerCanvas, render_mode='univariate')
    vis.scale = 20.0
    for mode in MODES:
        vis.render_mode = mode
        assert vis.render_mode == mode
        check_outputs(vis, mode)
    assert vis.scale == 20.0            # the camera survives a mode switch


def test_invalid_mode():
    vis = topsy.test(100, render_resolution=50, canvas_class=offscreen.VisualizerCanvas)
    with pytest.raises(ValueError, match="Invalid render_mode 'invalid'"):
        vis.render_mode = 'invalid'
    assert vis.render_mode == 'univariate'


@pytest.mark.parametrize("mode", MODES)
def test_mode_at_construction(mode):
    vis = topsy.test(100, render_resolution=50, canvas_class=offscreen.VisualizerCanvas, render_mode=mode)
    assert vis.render_mode == mode
    check_outputs(vis, mode)


class NoHdrCanvas(offscreen.VisualizerCanvas):
    def _rc_get_present_methods(self):
        return {"bitmap": {"formats": ["rgba-u8"]}}


def test_failed_switch_reverts():
    vis = topsy.test(100, render_resolution=50, canvas_class=NoHdrCanvas, render_mode='univariate')
    with pytest.raises(ValueError):
        vis.render_mode = 'rgb-hdr'
    assert vis.render_mode == 'univariate'
    check_outputs(vis, 'univariate')
    vis.render_mode = 'surface'           # a mode this canvas does support still switches afterwards
    check_outputs(vis, 'surface')


def test_unknown_quantity():
    vis = topsy.test(100, render_resolution=50, canvas_class=offscreen.VisualizerCanvas)
    with pytest.raises(ValueError):
        vis.quantity_name = "no-such-quantity"
    assert vis.quantity_name is None
