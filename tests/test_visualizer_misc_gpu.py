"""More of the drop-in surface on the GPU: split buffers, canvas interaction, array loader, save."""
import numpy as np
import numpy.testing as npt
import pytest

pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import topsy_b200 as topsy
from topsy_b200 import config, loader, visualizer
from topsy_b200.canvas import offscreen
from topsy_b200.drawreason import DrawReason


def test_split_buffers_give_the_same_image(monkeypatch):
    """N > MAX_PARTICLES_PER_BUFFER: several physical buffers, ranges split across them (split_buffers.py:78-116)."""
    one = topsy.test(5000, render_resolution=128, canvas_class=offscreen.VisualizerCanvas, with_cells=True)
    one.scale = 30.0
    one.render_sph(DrawReason.EXPORT)
    want = one._sph.get_image()
    monkeypatch.setattr(config, "MAX_PARTICLES_PER_BUFFER", 1300)
    many = topsy.test(5000, render_resolution=128, canvas_class=offscreen.VisualizerCanvas, with_cells=True)
    assert many.particle_buffers.num_buffers == 4
    many.scale = 30.0
    many.render_sph(DrawReason.EXPORT)
    got = many._sph.get_image()
    big = want[..., 0] > 1e-6 * want[..., 0].max()
    npt.assert_allclose(got[..., 0][big], want[..., 0][big], rtol=1e-4)
    # progressive (cell-mapped, multi-range) frames cross buffer boundaries correctly.  NB with cell mapping the reference
    # drops cells outside a sphere of 1.2 x scale (sph.py:313), so the fair comparison is progressive vs progressive.
    def progressive(vis):
        vis._sph._render_progression._recommended_num_particles_to_render = 700
        vis.render_sph(DrawReason.CHANGE)
        while vis._sph.needs_refine():
            vis.render_sph(DrawReason.REFINE)
        return vis._sph.get_image()

    p_many = progressive(many)
    p_one = progressive(one)
    npt.assert_allclose(p_many[..., 0][big], p_one[..., 0][big], rtol=1e-4)
    assert many._sph._engine.stats()["particles_submitted"] <= 5000


def test_create_and_write_split_buffers():
    """Mirrors the reference's test_create_buffers / test_write_buffers (/root/reference/tests/test_split_buffers.py:68-93):
    one device buffer per physical buffer, sized by its particle count; writes are checked for buffer and particle counts
    and land in the right slices."""
    from topsy_b200.device import Device
    from topsy_b200.split_buffers import SplitBuffers
    device = Device()
    sb = SplitBuffers(50, 15)
    buffers = sb.create_buffers(device, 4)
    assert len(buffers) == 4
    assert all(buf.numel() * buf.element_size() == 15 * 4 for buf in buffers[:-1])
    assert buffers[-1].numel() * buffers[-1].element_size() == 5 * 4
    data = np.arange(50, dtype=np.float32)
    with pytest.raises(ValueError):
        sb.write_buffers(device, buffers[:-1], data)        # wrong number of buffers
    with pytest.raises(ValueError):
        sb.write_buffers(device, buffers, data[:-1])        # wrong number of particles
    sb.write_buffers(device, buffers, data)
    for k, buf in enumerate(buffers):
        first, last = sb.buffer_range(k)
        npt.assert_array_equal(buf.cpu().numpy().view(np.float32), data[first:last])


def test_canvas_interaction():
    vis = topsy.test(2000, render_resolution=100, canvas_class=offscreen.VisualizerCanvas)
    vis.scale = 50.0
    c = vis.canvas
    c.submit_event({'event_type': 'resize', 'width': 400, 'height': 300, 'pixel_ratio': 1})
    assert (c.width_physical, c.height_physical) == (400, 300)
    c.submit_event({'event_type': 'pointer_move', 'x': 10, 'y': 10, 'buttons': [], 'modifiers': []})
    before = vis.rotation_matrix.copy()
    c.submit_event({'event_type': 'pointer_move', 'x': 40, 'y': 25, 'buttons': [1], 'modifiers': []})
    assert not np.allclose(vis.rotation_matrix, before)                      # drag rotates
    npt.assert_allclose(vis.rotation_matrix @ vis.rotation_matrix.T, np.eye(3), atol=1e-12)
    off = vis.position_offset.copy()
    c.submit_event({'event_type': 'pointer_move', 'x': 60, 'y': 25, 'buttons': [1], 'modifiers': ['Shift']})
    assert not np.allclose(vis.position_offset, off) and vis.crosshairs_visible   # shift-drag pans
    c.submit_event({'event_type': 'pointer_up'})
    assert not vis.crosshairs_visible
    c.submit_event({'event_type': 'wheel', 'dx': 0, 'dy': 1000})
    assert vis.scale == pytest.approx(50.0 * np.e)
    frame = c.draw()
    assert frame.shape == (300, 400, 4) and frame.dtype == np.uint8
    c.submit_event({'event_type': 'double_click', 'x': 200, 'y': 150})        # centre click: only the depth changes
    for _ in range(200):
        c.draw()
        if not c._later:
            break
    assert np.isfinite(vis.position_offset).all()
    c.submit_event({'event_type': 'key_up', 'key': 'r'})                      # autorange
    c.submit_event({'event_type': 'key_up', 'key': 'h'})                      # home view
    assert vis.scale == vis.data_loader.get_initial_view_width()


def test_array_loader_and_save(tmp_path):
    rs = np.random.RandomState(2)
    pos = rs.normal(size=(20000, 3)) * [3, 2, 1]
    vis = visualizer.Visualizer(data_loader_class=loader.ArrayDataLoader,
                                data_loader_args=(pos, np.full(20000, 0.2), np.ones(20000)),
                                data_loader_kwargs={"quantities": {"temp": np.exp(rs.normal(size=20000))},
                                                    "rgb": np.abs(rs.normal(size=(20000, 3)))},
                                render_resolution=128, canvas_class=offscreen.VisualizerCanvas)
    assert hasattr(vis.data_loader, "_cell_layout")
    assert vis.data_loader._cell_layout.get_num_cells() == config.DEFAULT_CELLS_NSIDE ** 3
    vis.quantity_name = "temp"
    img = vis.get_sph_image()
    assert img.shape == (128, 128) and np.nanmedian(img) > 0
    vis.save(str(tmp_path / "out.npy"))
    npt.assert_allclose(np.load(tmp_path / "out.npy"), vis.get_sph_image(), rtol=1e-5, equal_nan=True)
    vis.save(str(tmp_path / "out.png"))
    assert (tmp_path / "out.png").stat().st_size > 100
    vis.render_mode = 'rgb'
    assert vis.get_sph_image().shape == (128, 128, 3)
    depth = vis.get_depth_image()
    assert depth.shape == (128, 128)


def test_two_visualizers_share_the_device_without_interfering():
    a = topsy.test(3000, render_resolution=100, canvas_class=offscreen.VisualizerCanvas)
    b = topsy.test(3000, render_resolution=100, canvas_class=offscreen.VisualizerCanvas)
    assert a.device is b.device
    a.scale = 20.0
    b.scale = 100.0
    ia = a.get_sph_image().copy()
    ib = b.get_sph_image().copy()
    npt.assert_allclose(a.get_sph_image(), ia, rtol=1e-5)        # b's render did not clobber a's image or camera
    assert not np.allclose(ia, ib)
