"""Surface presentation: bilateral smoothing of the depth map, then lighting (reference: src/topsy/colormap/surface.py,
shaders/smooth.wgsl, shaders/surface.wgsl).  Same class name, parameters and methods; the compute + render passes become
the kernels behind ``tsplat_bilateral_filter`` (K10) and ``tsplat_surface_shade`` (K11)."""
from __future__ import annotations

import numpy as np
import torch

from .. import _native as N
from .. import config
from ..device import Texture
from .implementation import _OUT_FORMATS, Colormap


class ColorAsSurfaceMap(Colormap):
    """A colormap that renders surfaces with lighting instead of colormaps."""
    _default_params = {
        'depth_scale': 1.0,
        'light_direction': [0.0, 1.0 / np.sqrt(2.), 1.0 / np.sqrt(2.)],
        'light_color': [1.0, 1.0, 1.0],
        'ambient_color': [0.0, 0.0, 0.2],
        'smoothing_scale': 0.01,
        'weighted_average': False,
        'vmin': 0.0,
        'vmax': 1.0,
        'log': False,
        'colormap_name': config.DEFAULT_COLORMAP,
    }

    def __init__(self, device, input_texture, output_format, params):
        super().__init__(device, input_texture, output_format, params)
        self._surface_params = N.SurfaceParams()
        self._smoothed = None

    @classmethod
    def accepts_parameters(cls, parameters: dict) -> bool:
        return parameters.get("type", None) == "surface"

    # -- smoothing ------------------------------------------------------------------------------------------
    def _bilateral_parameters(self, width):
        """(spatial_sigma, range_sigma, kernel_size) as _encode_smoothing_filter_pass derives them (surface.py:267-280)."""
        sig = self._params.get('smoothing_scale', 0.01)
        if sig < 1e-5:
            sig = 1e-5
        spatial = np.float32(sig * width)
        n_pix = min(int(spatial * 4) + 1, config.MAX_SURFACE_SMOOTH_PIXELS)
        return spatial, np.float32(sig * 2), n_pix

    def _smooth(self, image: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        if out is None:
            if self._smoothed is None or self._smoothed.shape != image.shape or self._smoothed.device != image.device:
                self._smoothed = torch.empty_like(image)
            out = self._smoothed
        spatial, rng, n_pix = self._bilateral_parameters(image.shape[1])
        engine = self._device.engine(self._input_texture.tensor.shape[0])
        return engine.bilateral_filter(image, out, float(spatial), float(rng), n_pix)

    def _smooth_numpy(self, input_array: np.ndarray) -> np.ndarray:
        """Bilateral-filter a (height, width, 2) array on the device and read it back (surface.py:299-366)."""
        arr = np.asarray(input_array, dtype=np.float32)
        if arr.ndim != 3 or arr.shape[2] != 2:
            raise ValueError("Input array must be 3D with shape (height, width, 2)")
        source = self._device.upload(np.ascontiguousarray(arr))
        return self._smooth(source, torch.empty_like(source)).cpu().numpy()

    def sph_raw_output_to_content(self, numpy_image: np.ndarray):
        return self._smooth_numpy(numpy_image)

    # -- ranges ---------------------------------------------------------------------------------------------
    def autorange_vmin_vmax(self, vals):
        """Range of the material value over the pixels that received a fragment (surface.py:256-259)."""
        valid = vals[..., 1].ravel() > 0.0
        self._autorange_using_values(vals[..., 0].ravel()[valid])

    def autorange_device(self, image: torch.Tensor, mass_scale: float):
        """Same decisions as ``autorange_vmin_vmax`` without the read-back.  The surface image holds maxima, not sums, so
        there is no mass scale."""
        self._autorange_device_values(image, N.CONTENT_CH0_WHERE_CH1, 1.0)

    def _update_parameter_buffer(self, width, height, mass_scale):
        p = self._surface_params
        p.depth_scale = self._params.get('depth_scale', 1.0)
        p.light_direction[:] = [float(v) for v in self._params.get('light_direction', [0.0, 0.0, 1.0])]
        p.light_color[:] = [float(v) for v in self._params.get('light_color', [1.0, 1.0, 1.0])]
        p.ambient_color[:] = [float(v) for v in self._params.get('ambient_color', [0.2, 0.2, 0.2])]
        p.window_aspect_ratio = float(width) / height
        p.vmin = np.float32(self.get_parameter("vmin"))
        p.vmax = np.float32(self.get_parameter("vmax"))

    # -- presentation ---------------------------------------------------------------------------------------
    def _launch(self, image: torch.Tensor, target: Texture):
        fmt, _, _ = _OUT_FORMATS[target.format]
        p = self._surface_params
        p.material_colormap = int(bool(self.get_parameter('weighted_average')))       # MATERIAL_COLORMAP
        p.log_scale = int(bool(self.get_parameter('weighted_average')) and bool(self.get_parameter('log')))
        smoothed = self._smooth(image)
        lut = self._texture if p.material_colormap else None
        self._device.engine(image.shape[0]).surface_shade(smoothed, p, lut, target.tensor, fmt)
