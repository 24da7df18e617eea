"""Colormap stage on the GPU against the numpy restatement of colormap.wgsl (oracle) -- the reference's
tests/test_colormap.py compares against matplotlib's software colormap with atol 5/255; same idea, tighter bound."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import topsy_oracle as o
from topsy_b200 import colormap as cm
from topsy_b200.colormap import luts
from topsy_b200.device import Device


@pytest.fixture(scope="module")
def device():
    return Device()


def make_image(R=96, seed=0, signed=False):
    rs = np.random.RandomState(seed)
    den = np.exp(rs.normal(size=(R, R)) * 2.0).astype(np.float32)
    den[rs.uniform(size=(R, R)) < 0.02] = 0.0                       # empty pixels -> log(0), 0/0
    q = rs.normal(size=(R, R)).astype(np.float32) if signed else np.exp(rs.normal(size=(R, R))).astype(np.float32)
    return np.stack([den, den * q], axis=-1)


def holder_for(device, image, params, fmt="rgba8unorm"):
    tex = device.create_texture((image.shape[1], image.shape[0], 1), "rg32float" if image.shape[2] == 2 else "rgba32float")
    tex.tensor.copy_(torch.from_numpy(image))
    h = cm.ColormapHolder(device, tex, fmt)
    h.update_parameters(params)
    return h, tex


def close_u8(a, b, frac=2e-3):
    d = np.abs(a.astype(int) - b.astype(int))
    assert d.max() <= 1, d.max()
    assert (d > 0).mean() <= frac * 4


@pytest.mark.parametrize("weighted,log", [(False, True), (False, False), (True, True), (True, False)])
def test_scalar_maps(device, weighted, log):
    img = make_image(signed=False)
    h, tex = holder_for(device, img, {'type': 'density', 'colormap_name': 'viridis', 'weighted_average': weighted})
    h.autorange(img)
    h.update_parameters({'log': log})
    mass_scale = 3.0
    h.set_scaling(96, 96, mass_scale)
    out = device.create_texture((96, 96, 1), "rgba8unorm")
    h.encode_render_pass(None, out)
    p = o.colormap_params(h['vmin'], h['vmax'], log, weighted, mass_scale)
    want = o.to_unorm8(o.colormap_scalar(img, p, luts.colormap_table_1d('viridis', 1000), log, weighted))
    close_u8(out.tensor.cpu().numpy(), want)


def test_autorange_matches_oracle(device):
    img = make_image(signed=True)
    h, _ = holder_for(device, img, {'type': 'density', 'colormap_name': 'viridis', 'weighted_average': True})
    h.autorange(img)
    want = o.autorange_scalar(img[..., 1] / img[..., 0])
    assert h['log'] == want['log'] is False
    assert h['vmin'] == pytest.approx(want['vmin']) and h['vmax'] == pytest.approx(want['vmax'])


def test_bivariate_map(device):
    img = make_image(signed=True, seed=3)
    h, tex = holder_for(device, img, {'type': 'bivariate', 'colormap_name': 'twilight_shifted', 'weighted_average': True})
    assert isinstance(h._impl, cm.BivariateColormap)
    h.autorange(img)
    h.set_scaling(96, 96, 1.0)
    out = device.create_texture((96, 96, 1), "rgba8unorm")
    h.encode_render_pass(None, out)
    p = o.colormap_params(h['vmin'], h['vmax'], h['log'], True, 1.0, density_vmin=h['density_vmin'], density_vmax=h['density_vmax'])
    want = o.to_unorm8(o.colormap_bivariate(img, p, luts.colormap_table_2d('twilight_shifted', 1000), h['log'], True))
    close_u8(out.tensor.cpu().numpy(), want, frac=5e-3)


@pytest.mark.parametrize("hdr", [False, True])
def test_rgb_maps(device, hdr):
    rs = np.random.RandomState(5)
    img = np.exp(rs.normal(size=(64, 64, 4)) * 2).astype(np.float32)
    img[..., 3] = rs.randint(0, 30, (64, 64))
    fmt = "rgba16float" if hdr else "rgba8unorm"
    h, tex = holder_for(device, img, {'type': 'rgb', 'hdr': hdr, 'log': True}, fmt)
    assert type(h._impl) is (cm.RGBHDRColormap if hdr else cm.RGBColormap)
    h.autorange(img)
    want_range = o.autorange_rgb(img, 99.0 if hdr else 99.9, 2.5 if hdr else 3.0)
    assert h['vmax'] == pytest.approx(want_range['vmax']) and h['vmin'] == pytest.approx(want_range['vmin'])
    h.update_parameters({'gamma': 0.7})
    h.set_scaling(64, 64, 2.0)
    out = device.create_texture((64, 64, 1), fmt)
    h.encode_render_pass(None, out)
    p = o.colormap_params(h['vmin'], h['vmax'], True, False, 2.0, gamma=0.7, may_produce_weighted_average=False)
    want = o.colormap_rgb(img, p)
    got = out.tensor.cpu().numpy()
    if hdr:
        np.testing.assert_allclose(got.astype(np.float32), want.astype(np.float16).astype(np.float32), rtol=2e-3, atol=1e-4)
    else:
        close_u8(got, o.to_unorm8(want))


def test_mag_parameters(device):
    img = np.ones((8, 8, 4), np.float32)
    h, _ = holder_for(device, img, {'type': 'rgb', 'hdr': False, 'log': True})
    h.update_parameters({'min_mag': 20.0, 'max_mag': 25.0})
    assert h['min_mag'] == pytest.approx(20.0) and h['max_mag'] == pytest.approx(25.0)
    assert h['vmax'] == pytest.approx(o.mag_per_arcsec2_to_log_output(20.0))
    assert h['vmin'] == pytest.approx(o.mag_per_arcsec2_to_log_output(25.0))


def test_mag_parameters_known_answers(device):
    """The reference's own known answers for the vmin/vmax <-> magnitude conversion
    (/root/reference/tests/test_colormap.py:88-104), in both directions through the holder."""
    img = np.ones((8, 8, 4), np.float32)
    h, _ = holder_for(device, img, {'type': 'rgb', 'hdr': False, 'log': True})
    h.update_parameters({'vmin': 1.0, 'vmax': 2.0})
    assert h.get_parameter('vmin') == 1.0 and h.get_parameter('vmax') == 2.0
    assert np.allclose(h.get_parameter('min_mag'), 31.57212566586528)
    assert np.allclose(h.get_parameter('max_mag'), 34.07212566586528)
    h.update_parameters({'min_mag': 1.0, 'max_mag': 2.0})
    assert np.allclose(h.get_parameter('min_mag'), 1.0) and np.allclose(h.get_parameter('max_mag'), 2.0)
    assert np.allclose(h.get_parameter('vmin'), 13.828850266346112)
    assert np.allclose(h.get_parameter('vmax'), 14.228850266346113)
    h['vmin'] = 5.0                                   # dictionary access (:107-117)
    assert h['vmin'] == 5.0 and h['type'] == 'rgb'


def test_holder_dispatch_and_host_roundtrip(device):
    img = make_image(R=32)
    h, _ = holder_for(device, img, {'type': 'density', 'colormap_name': 'viridis'})
    assert type(h._impl) is cm.Colormap
    assert h.update_parameters({'type': 'bivariate'}) is True
    assert type(h._impl) is cm.BivariateColormap
    assert h.update_parameters({'vmin': 0.0}) is False
    with pytest.raises(ValueError):
        cm.ColormapHolder(device, None, "rgba8unorm").autorange(img)        # not initialised yet
    h.update_parameters({'type': 'density', 'vmin': -1.0, 'vmax': 1.0, 'log': True})
    rgba = h.sph_raw_output_to_image(img)
    assert rgba.shape == (32, 32, 4) and rgba.dtype == np.uint8
    p = o.colormap_params(-1.0, 1.0, True, False, 1.0)
    close_u8(rgba, o.to_unorm8(o.colormap_scalar(img, p, luts.colormap_table_1d('viridis', 1000), True, False)), frac=5e-3)
    with pytest.raises(ValueError):
        h.sph_raw_output_to_image(img.astype(np.float64))


def test_presentation_resampling(device):
    """Non-square / non-native output: the square image covers the larger window dimension (colormap.wgsl:41-73)."""
    R = 64
    img = np.zeros((R, R, 2), np.float32)
    img[..., 0] = np.linspace(1.0, 100.0, R)[None, :]            # varies along x only
    h, _ = holder_for(device, img, {'type': 'density', 'colormap_name': 'gray', 'vmin': 0.0, 'vmax': 2.0, 'log': True})
    out = device.create_texture((128, 64, 1), "rgba8unorm")       # 2:1 window
    h.set_scaling(128, 64, 1.0)
    h.encode_render_pass(None, out)
    got = out.tensor.cpu().numpy()
    assert got.shape == (64, 128, 4)
    assert (np.diff(got[32, :, 0].astype(int)) >= 0).all() and got[32, -1, 0] > got[32, 0, 0] + 100
    assert np.abs(got[10].astype(int) - got[50].astype(int)).max() <= 1     # rows identical: y is cropped, not stretched


@pytest.mark.parametrize("kind", ["density", "weighted", "signed", "bivariate", "rgb", "rgb-hdr", "tiny"])
def test_device_autorange_equals_host_autorange(device, kind):
    """K8 (stats + radix select on the device) against the reference's host rule (numpy percentiles of the read-back)."""
    rs = np.random.RandomState(11)
    R = 160 if kind != "tiny" else 12
    if kind.startswith("rgb"):
        img = np.exp(rs.normal(size=(R, R, 4)) * 2).astype(np.float32)
        img[..., 3] = rs.randint(0, 30, (R, R))
        img[rs.uniform(size=(R, R)) < 0.05] = 0.0
        params = {'type': 'rgb', 'hdr': kind == "rgb-hdr", 'log': True}
        fmt = "rgba16float" if kind == "rgb-hdr" else "rgba8unorm"
    else:
        img = make_image(R=R, seed=21, signed=(kind in ("signed", "bivariate")))
        params = {'type': 'bivariate' if kind == "bivariate" else 'density', 'colormap_name': 'viridis',
                  'weighted_average': kind in ("weighted", "signed", "bivariate")}
        fmt = "rgba8unorm"
    scale = 2.5
    host, _ = holder_for(device, img, params, fmt)
    dev, _ = holder_for(device, img, params, fmt)
    host.autorange(img * np.float32(scale))
    dev.autorange_texture(scale)
    for key in ('vmin', 'vmax', 'log', 'density_vmin', 'density_vmax'):
        a, b = host[key], dev[key]
        if a is None or isinstance(a, (bool, np.bool_)):
            assert a == b, key
        else:
            assert b == pytest.approx(a, rel=2e-5, abs=2e-5), key
    for key in ('ui_range_linear', 'ui_range_log', 'ui_range_density'):
        if host[key] is not None:
            np.testing.assert_allclose(dev[key], host[key], rtol=2e-5, atol=2e-5)
