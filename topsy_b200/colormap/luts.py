"""Colormap tables: ``name -> RGBA float32`` samples, the data the reference takes from matplotlib
(src/topsy/colormap/implementation.py:235-238 ``cmap(np.linspace(0.001, 0.999, N))`` and :590-606 for the 2-D map).

matplotlib is an optional dependency here.  Resolution order:
  1. matplotlib, if importable (identical to the reference);
  2. OpenCV's 256-entry tables (cv2 ships matplotlib's viridis/magma/inferno/plasma/cividis/twilight/
     twilight_shifted/turbo and the classic jet/hot/... maps), evaluated with ListedColormap semantics;
  3. a small set of analytic maps (gray, and cubehelix under the requested name) so the pipeline always runs.
The GPU colormap kernel takes any (N,4) / (N,N,4) float32 table, so user-supplied tables work too (register_table).
"""
from __future__ import annotations

import logging

import numpy as np

logger = logging.getLogger(__name__)

_USER_TABLES: dict[str, np.ndarray] = {}

_CV2_NAMES = {"viridis": "VIRIDIS", "magma": "MAGMA", "inferno": "INFERNO", "plasma": "PLASMA", "cividis": "CIVIDIS",
              "twilight": "TWILIGHT", "twilight_shifted": "TWILIGHT_SHIFTED", "turbo": "TURBO", "jet": "JET", "hot": "HOT",
              "bone": "BONE", "cool": "COOL", "spring": "SPRING", "summer": "SUMMER", "autumn": "AUTUMN", "winter": "WINTER",
              "pink": "PINK", "ocean": "OCEAN", "rainbow": "RAINBOW", "hsv": "HSV"}


def register_table(name: str, rgba: np.ndarray):
    """Make an (M,3|4) table available under ``name`` (values in [0,1])."""
    t = np.asarray(rgba, dtype=np.float64)
    if t.ndim != 2 or t.shape[1] not in (3, 4):
        raise ValueError("table must be (M,3) or (M,4)")
    if t.shape[1] == 3:
        t = np.concatenate([t, np.ones((len(t), 1))], axis=1)
    _USER_TABLES[name] = t


def _listed_lookup(table: np.ndarray, x: np.ndarray) -> np.ndarray:
    """matplotlib ListedColormap.__call__ for floats in [0,1]: index int(x*N), clipped."""
    n = len(table)
    idx = np.clip((np.asarray(x, np.float64) * n).astype(np.int64), 0, n - 1)
    return table[idx]


def _cubehelix(x, start=0.5, rot=-1.5, hue=1.0):
    x = np.asarray(x, np.float64)
    ang = 2 * np.pi * (start / 3.0 + 1.0 + rot * x)
    amp = hue * x * (1 - x) / 2.0
    r = x + amp * (-0.14861 * np.cos(ang) + 1.78277 * np.sin(ang))
    g = x + amp * (-0.29227 * np.cos(ang) - 0.90649 * np.sin(ang))
    b = x + amp * (1.97294 * np.cos(ang))
    return np.clip(np.stack([r, g, b, np.ones_like(x)], axis=-1), 0, 1)


def sample_colormap(name: str, x: np.ndarray) -> np.ndarray:
    """RGBA (float64, shape x.shape + (4,)) of colormap ``name`` at positions x in [0,1]."""
    x = np.asarray(x, np.float64)
    if name in _USER_TABLES:
        return _listed_lookup(_USER_TABLES[name], x)
    try:
        import matplotlib
        return np.asarray(matplotlib.colormaps[name](x), dtype=np.float64)
    except ImportError:
        pass
    reverse = name.endswith("_r")
    base = name[:-2] if reverse else name
    if reverse:
        x = 1.0 - x
    if base in ("gray", "grey", "Greys_r"):
        return np.stack([x, x, x, np.ones_like(x)], axis=-1)
    if base in _CV2_NAMES:
        try:
            import cv2
            ramp = np.arange(256, dtype=np.uint8).reshape(1, 256)
            bgr = cv2.applyColorMap(ramp, getattr(cv2, "COLORMAP_" + _CV2_NAMES[base]))[0]
            table = np.concatenate([bgr[:, ::-1].astype(np.float64) / 255.0, np.ones((256, 1))], axis=1)
            return _listed_lookup(table, x)
        except ImportError:
            pass
    logger.warning("colormap '%s' unavailable without matplotlib/cv2; using cubehelix", name)
    return _cubehelix(x)


def colormap_table_1d(name: str, num_points: int) -> np.ndarray:
    """(num_points, 4) float32 -- Colormap._generate_mapping_rgba_f32 (implementation.py:235-238)."""
    return sample_colormap(name, np.linspace(0.001, 0.999, num_points)).astype(np.float32)


def _rgb_to_hsv(rgb):
    out = np.empty_like(rgb)
    flat_in = rgb.reshape(-1, 3)
    flat_out = out.reshape(-1, 3)
    # vectorised version of colorsys.rgb_to_hsv / matplotlib.colors.rgb_to_hsv
    mx = flat_in.max(axis=1); mn = flat_in.min(axis=1)
    delta = mx - mn
    s = np.where(mx > 0, delta / np.where(mx > 0, mx, 1), 0.0)
    h = np.zeros_like(mx)
    nz = delta > 0
    r, g, b = flat_in[:, 0], flat_in[:, 1], flat_in[:, 2]
    d = np.where(nz, delta, 1)
    idx = nz & (r == mx)
    h[idx] = ((g - b) / d)[idx]
    idx = nz & (g == mx) & ~(r == mx)
    h[idx] = 2.0 + ((b - r) / d)[idx]
    idx = nz & (b == mx) & ~(r == mx) & ~(g == mx)
    h[idx] = 4.0 + ((r - g) / d)[idx]
    h = (h / 6.0) % 1.0
    flat_out[:, 0] = h; flat_out[:, 1] = s; flat_out[:, 2] = mx
    return out


def _hsv_to_rgb(hsv):
    h, s, v = hsv[..., 0], hsv[..., 1], hsv[..., 2]
    i = (h * 6.0).astype(np.int64)
    f = h * 6.0 - i
    p = v * (1 - s); q = v * (1 - s * f); t = v * (1 - s * (1 - f))
    i = i % 6
    r = np.choose(i, [v, q, p, p, t, v]); g = np.choose(i, [t, v, v, q, p, p]); b = np.choose(i, [p, p, t, v, v, q])
    gray = s == 0
    r = np.where(gray, v, r); g = np.where(gray, v, g); b = np.where(gray, v, b)
    return np.stack([r, g, b], axis=-1)


def colormap_table_2d(name: str, num_points: int) -> np.ndarray:
    """(num_points, num_points, 4) float32 bivariate table: hue from the colormap along axis 0 (the quantity),
    HSV value ramp along axis 1 (the density), saturation faded out over the brightest quarter
    (BivariateColormap._generate_mapping_rgba_f32, implementation.py:590-606)."""
    ramp = np.linspace(0.001, 0.999, num_points)
    rgba = np.ones((num_points, num_points, 4), dtype=np.float32)
    rgba[:, :, :] = sample_colormap(name, ramp)[:, np.newaxis, :]
    hsv = _rgb_to_hsv(rgba[..., :3].astype(np.float64))
    hsv[..., 2] = ramp[np.newaxis, :]
    fade = np.ones(num_points)
    fade[3 * num_points // 4:] = np.linspace(1.0, 0.0, num_points // 4)
    hsv[..., 1] *= fade[np.newaxis, :]
    rgba[..., :3] = _hsv_to_rgb(hsv)
    return rgba
