"""Presentation stage: accumulation image -> RGBA (reference: src/topsy/colormap/implementation.py + shaders/colormap.wgsl).

Class names, parameter dictionaries, autorange rules and the mass-scale correction follow the reference; the full-screen
quad + fragment shader is replaced by the fused normalise / log / LUT kernel behind ``tsplat_colormap``.  The shader's
compile-time switches (WEIGHTED_MEAN, LOG_SCALE, BIVARIATE) become the ``kind`` / ``log_scale`` fields of the kernel's
parameter block, so changing them never rebuilds anything.
"""
from __future__ import annotations

import logging

import numpy as np
import torch

from .. import _native as N
from .. import config
from ..device import Texture
from . import luts

logger = logging.getLogger(__name__)

_OUT_FORMATS = {"rgba8unorm": (N.FMT_RGBA8, torch.uint8, np.uint8), "rgba16float": (N.FMT_RGBA16F, torch.float16, np.float16),
                "rgba32float": (N.FMT_RGBA32F, torch.float32, np.float32)}


class ColormapBase:
    _default_params = {}

    def __init__(self, device, input_texture, output_format, params: dict):
        self._device = device
        self._input_texture = input_texture
        self._output_format = output_format
        self._params = self._default_params | params

    @classmethod
    def accepts_parameters(cls, parameters: dict) -> bool:
        return False

    def update_parameters(self, parameters: dict):
        if not self.accepts_parameters(self._params | parameters):
            raise ValueError(f"Colormap {self.__class__.__name__} does not accept parameter update: {parameters}")
        self._params.update(parameters)

    def get_parameter(self, name: str):
        return self._params.get(name, None)

    def get_parameters(self) -> dict:
        return self._params.copy()

    def encode_render_pass(self, command_encoder, target_texture_view, bind_group=None):
        raise NotImplementedError("Subclasses must implement encode_render_pass")

    def set_scaling(self, output_width, output_height, mass_scaling):
        raise NotImplementedError("Subclasses must implement set_scaling")


class NoColormap(ColormapBase):
    """Placeholder until the visualizer has decided which map it needs."""

    @classmethod
    def accepts_parameters(cls, parameters: dict) -> bool:
        return parameters.get("type", None) == "none"


class Colormap(ColormapBase):
    """Scalar map: density or mass-weighted mean, linear or log10, through a 1-D LUT (colormap.wgsl:109-127)."""
    input_channels = 2
    percentile_scaling = [1.0, 99.9]
    may_produce_weighted_average = True
    _default_params = {'colormap_name': 'viridis', 'vmin': 0.0, 'vmax': 1.0, 'log': True, 'weighted_average': False}

    def __init__(self, device, input_texture, output_format, params):
        super().__init__(device, input_texture, output_format, params)
        self._kernel_params = N.ColormapParams()
        self._setup_map_texture()

    @classmethod
    def accepts_parameters(cls, parameters: dict) -> bool:
        return parameters.get("type", None) == "density"

    def update_parameters(self, parameters: dict):
        before = self.get_parameter('colormap_name')
        super().update_parameters(parameters)
        if self.get_parameter('colormap_name') != before:
            self._setup_map_texture()

    # -- LUT ------------------------------------------------------------------------------------------------
    def _generate_mapping_rgba_f32(self, num_points):
        return luts.colormap_table_1d(self._params.get('colormap_name', config.DEFAULT_COLORMAP), num_points)

    def _setup_map_texture(self, num_points=config.COLORMAP_NUM_SAMPLES):
        rgba = np.ascontiguousarray(self._generate_mapping_rgba_f32(num_points), dtype=np.float32)
        self._texture = self._device.upload(rgba)

    # -- image -> logical content / RGBA ----------------------------------------------------------------------
    def sph_raw_output_to_content(self, numpy_image: np.ndarray):
        """Density -> channel 0; weighted mean -> channel 1 / channel 0 (implementation.py:119-130)."""
        if self._params['weighted_average']:
            with np.errstate(divide='ignore', invalid='ignore'):
                return numpy_image[..., 1] / numpy_image[..., 0]
        return numpy_image[..., 0]

    def _kernel_kind(self):
        return N.CMAP_WEIGHTED if self._params.get('weighted_average', False) else N.CMAP_DENSITY

    def _launch(self, image: torch.Tensor, target: Texture):
        fmt, _, _ = _OUT_FORMATS[target.format]
        self._kernel_params.kind = self._kernel_kind()
        self._kernel_params.log_scale = int(bool(self._params['log']))
        lut = self._texture if self._kernel_params.kind != N.CMAP_RGB else None
        self._device.engine(image.shape[0]).colormap(image, self._kernel_params, lut, target.tensor, fmt)

    def encode_render_pass(self, command_encoder, target_texture_view, bind_group=None):
        """Run the colormap kernel on the bound SPH image (or on ``bind_group``, an alternative input Texture) into
        ``target_texture_view``.  ``command_encoder`` is accepted for signature compatibility and ignored."""
        source = bind_group if bind_group is not None else self._input_texture
        self._launch(source.tensor, target_texture_view)

    def sph_raw_output_to_image(self, numpy_image: np.ndarray):
        """Host image in, colormapped host image out (implementation.py:132-201)."""
        if numpy_image.ndim != 3:
            raise ValueError(f"Expected a 3D array, but got shape {numpy_image.shape}")
        if numpy_image.shape[2] != self.input_channels:
            raise ValueError(f"Expected the last dimension to have size {self.input_channels}, but got {numpy_image.shape[2]}")
        if numpy_image.dtype != np.float32:
            raise ValueError(f"Expected dtype to be np.float32, but got {numpy_image.dtype}")
        if self._output_format not in ("rgba8unorm", "rgba32float"):
            raise ValueError(f"Unsupported output format: {self._output_format}")
        if numpy_image.shape[0] != numpy_image.shape[1]:
            raise ValueError("Expected a square image")
        channels = 4 if numpy_image.shape[2] == 3 else numpy_image.shape[2]
        padded = np.zeros(numpy_image.shape[:2] + (channels,), np.float32)
        padded[..., :numpy_image.shape[2]] = numpy_image
        source = self._device.upload(padded)
        target = self._device.create_texture((numpy_image.shape[1], numpy_image.shape[0], 1), self._output_format)
        self.set_scaling(numpy_image.shape[1], numpy_image.shape[0], 1.0)
        self._launch(source, target)
        return target.tensor.cpu().numpy()

    # -- ranges ---------------------------------------------------------------------------------------------
    def set_scaling(self, width, height, scaling):
        self._update_parameter_buffer(width, height, scaling)

    @classmethod
    def _finite_range(cls, values):
        good = values[np.isfinite(values)]
        return (np.min(good), np.max(good)) if len(good) > 0 else (np.nan, np.nan)

    def autorange_vmin_vmax(self, vals):
        """Pick vmin/vmax (and log vs linear) from the most recent SPH image (implementation.py:381-425)."""
        self._autorange_using_values(self.sph_raw_output_to_content(vals).ravel())

    def _autorange_using_values(self, vals):
        with np.errstate(divide='ignore', invalid='ignore'):
            log_lo, log_hi = self._finite_range(np.log10(vals))
            lin_lo, lin_hi = self._finite_range(vals)
            if log_hi == log_lo:
                log_hi += 1.0; log_lo -= 1.0
            if lin_hi == lin_lo:
                lin_hi += 1.0; lin_lo -= 1.0
            new_params = {'ui_range_linear': (lin_lo, lin_hi), 'ui_range_log': (log_lo, log_hi),
                          'log': not (vals < 0).any()}
            if new_params['log']:
                vals = np.log10(vals)
        vals = vals[np.isfinite(vals)]
        if len(vals) > 200:
            self._params['vmin'], self._params['vmax'] = np.percentile(vals, self.percentile_scaling)
        elif len(vals) > 2:
            self._params['vmin'], self._params['vmax'] = np.min(vals), np.max(vals)
        else:
            logger.warning("Problem setting vmin/vmax, perhaps there are no particles or something is wrong with them?")
            self._params['vmin'], self._params['vmax'] = 0.0, 1.0
        self.update_parameters(new_params)
        logger.info(f"Autoscale: log_scale={self._params['log']}, vmin={self._params['vmin']}, vmax={self._params['vmax']}")

    # -- the same decisions taken on the device (no read-back of the image) ------------------------------------------
    def _content_kind(self):
        return N.CONTENT_RATIO if self._params.get('weighted_average', False) else N.CONTENT_CH0

    def autorange_device(self, image: torch.Tensor, mass_scale: float):
        """``autorange_vmin_vmax(image.cpu() * mass_scale)`` without the read-back: statistics and exact order statistics
        come from the K8 kernels (tsplat_content_stats / tsplat_content_select)."""
        self._autorange_device_values(image, self._content_kind(), mass_scale)

    def _autorange_device_values(self, image, content, mass_scale):
        eng = self._device.engine(image.shape[0])
        st = eng.content_stats(image, content, mass_scale)
        lin_lo, lin_hi, log_lo, log_hi = st.lin_min, st.lin_max, st.log_min, st.log_max
        if log_hi == log_lo:
            log_hi += 1.0; log_lo -= 1.0
        if lin_hi == lin_lo:
            lin_hi += 1.0; lin_lo -= 1.0
        use_log = not bool(st.any_negative)
        n = st.n_finite_log if use_log else st.n_finite_lin
        if n > 200:
            vmin, vmax = eng.content_percentiles(image, content, mass_scale, use_log, n, self.percentile_scaling)
        elif n > 2:
            vmin, vmax = (st.log_min, st.log_max) if use_log else (st.lin_min, st.lin_max)
        else:
            logger.warning("Problem setting vmin/vmax, perhaps there are no particles or something is wrong with them?")
            vmin, vmax = 0.0, 1.0
        self._params['vmin'], self._params['vmax'] = vmin, vmax
        self.update_parameters({'ui_range_linear': (lin_lo, lin_hi), 'ui_range_log': (log_lo, log_hi), 'log': use_log})

    def _update_parameter_buffer(self, width, height, mass_scale):
        """The device image is *unscaled* (sum over the particles rendered so far), so the ranges are shifted instead
        (implementation.py:427-453).  A weighted mean is a ratio and needs no correction."""
        p = self._kernel_params
        d_vmin = self._params.get('density_vmin', 0.0)
        d_vmax = self._params.get('density_vmax', 1.0)
        d_vmin = 0.0 if d_vmin is None else d_vmin
        d_vmax = 1.0 if d_vmax is None else d_vmax
        p.density_vmin = np.float32(d_vmin - np.log10(mass_scale))
        p.density_vmax = np.float32(d_vmax - np.log10(mass_scale))
        if self.may_produce_weighted_average and self._params.get('weighted_average', False):
            mass_scale = 1.0
        vmin = np.float32(self._params['vmin']); vmax = np.float32(self._params['vmax'])
        if self._params['log']:
            vmin = np.float32(vmin - np.log10(mass_scale)); vmax = np.float32(vmax - np.log10(mass_scale))
        else:
            vmin = np.float32(vmin / mass_scale); vmax = np.float32(vmax / mass_scale)
        p.vmin, p.vmax = vmin, vmax
        p.window_aspect_ratio = float(width) / height
        p.gamma = self._params.get('gamma', 1.0)


class RGBColormap(Colormap):
    """Three-band log + gamma map without LUT, surface-brightness style ranges (colormap.wgsl:131-159)."""
    input_channels = 3
    max_percentile = 99.9
    dynamic_range = 3.0
    may_produce_weighted_average = False
    _sterrad_to_arcsec2 = 2.3504430539466191e-11
    _default_params = {'vmin': 0.0, 'vmax': 1.0, 'log': True, 'gamma': 1.0}

    @classmethod
    def accepts_parameters(cls, parameters: dict) -> bool:
        parameters = cls._default_params | parameters
        return parameters.get("type", None) == "rgb" and (not parameters['hdr']) and parameters['log']

    def _setup_map_texture(self, num_points=None):
        self._texture = None

    def _kernel_kind(self):
        return N.CMAP_RGB

    @classmethod
    def _log_output_to_mag_per_arcsec2(cls, val):
        return None if val is None else -2.5 * (val + np.log10(cls._sterrad_to_arcsec2) - 4)     # +4: (10 pc -> kpc)^2

    @classmethod
    def _mag_per_arcsec2_to_log_output(cls, val):
        return None if val is None else val / -2.5 + 4 - np.log10(cls._sterrad_to_arcsec2)

    def get_parameters(self) -> dict:
        params = super().get_parameters()
        params['min_mag'] = self._log_output_to_mag_per_arcsec2(params['vmax'])
        params['max_mag'] = self._log_output_to_mag_per_arcsec2(params['vmin'])
        return params

    def get_parameter(self, name: str):
        if name == "min_mag":
            return self._log_output_to_mag_per_arcsec2(self.get_parameter("vmax"))
        if name == "max_mag":
            return self._log_output_to_mag_per_arcsec2(self.get_parameter("vmin"))
        return super().get_parameter(name)

    def update_parameters(self, parameters: dict):
        parameters = dict(parameters)
        if "min_mag" in parameters:
            parameters['vmax'] = self._mag_per_arcsec2_to_log_output(parameters['min_mag'])
        if "max_mag" in parameters:
            parameters['vmin'] = self._mag_per_arcsec2_to_log_output(parameters['max_mag'])
        ColormapBase.update_parameters(self, parameters)

    def autorange_vmin_vmax(self, vals):
        """vmax = high percentile of log10 of every channel value, vmin a fixed dynamic range below (:512-531)."""
        with np.errstate(divide='ignore', invalid='ignore'):
            vals = np.log10(vals.ravel())
        vals = vals[np.isfinite(vals)]
        if len(vals) > 200:
            self._params['vmax'] = np.percentile(vals, self.max_percentile)
        elif len(vals) > 2:
            self._params['vmax'] = np.max(vals)
        else:
            logger.warning("Problem setting vmin/vmax, perhaps there are no particles or something is wrong with them?")
            self._params['vmax'] = 1.0
        self._params['vmin'] = self._params['vmax'] - self.dynamic_range

    def sph_raw_output_to_content(self, numpy_image: np.ndarray):
        return numpy_image[..., :3]

    def autorange_device(self, image: torch.Tensor, mass_scale: float):
        eng = self._device.engine(image.shape[0])
        st = eng.content_stats(image, N.CONTENT_ALL, mass_scale)
        n = st.n_finite_log
        if n > 200:
            self._params['vmax'] = eng.content_percentiles(image, N.CONTENT_ALL, mass_scale, True, n, [self.max_percentile])[0]
        elif n > 2:
            self._params['vmax'] = st.log_max
        else:
            self._params['vmax'] = 1.0
        self._params['vmin'] = self._params['vmax'] - self.dynamic_range


class RGBHDRColormap(RGBColormap):
    max_percentile = 99.0
    dynamic_range = 2.5      # the SDR-equivalent range; HDR output exceeds 1.0 above it

    @classmethod
    def accepts_parameters(cls, parameters: dict) -> bool:
        parameters = cls._default_params | parameters
        return parameters.get("type", None) == "rgb" and parameters['hdr'] and parameters['log']


class BivariateColormap(Colormap):
    """2-D LUT indexed by (log density, value) (colormap.wgsl:84-107)."""
    default_quantity_name = 'rho'
    _default_params = Colormap._default_params | {'density_vmin': 0.0, 'density_vmax': 1.0, 'ui_range_density': (0.0, 1.0)}

    @classmethod
    def accepts_parameters(cls, parameters: dict) -> bool:
        return parameters.get("type", None) == "bivariate" and (not parameters.get("hdr", False))

    def _kernel_kind(self):
        return N.CMAP_BIVARIATE_WEIGHTED if self._params.get('weighted_average', False) else N.CMAP_BIVARIATE

    def _generate_mapping_rgba_f32(self, num_points):
        return luts.colormap_table_2d(self._params['colormap_name'], num_points)

    def autorange_device(self, image: torch.Tensor, mass_scale: float):
        eng = self._device.engine(image.shape[0])
        st = eng.content_stats(image, N.CONTENT_CH0, mass_scale)
        density_vmin, density_vmax = eng.content_percentiles(image, N.CONTENT_CH0, mass_scale, True, st.n_finite_log,
                                                             self.percentile_scaling)
        self.update_parameters({'density_vmin': density_vmin, 'density_vmax': density_vmax,
                                'ui_range_density': (st.log_min, st.log_max)})
        self._autorange_device_values(image, self._content_kind(), mass_scale)

    def sph_raw_output_to_content(self, numpy_image: np.ndarray):
        out = numpy_image.copy()
        if self._params['weighted_average']:
            with np.errstate(divide='ignore', invalid='ignore'):
                out[..., 1] /= out[..., 0]
        else:
            out[..., 1] = out[..., 0]
        return out

    def autorange_vmin_vmax(self, vals):
        vals = self.sph_raw_output_to_content(vals)
        with np.errstate(divide='ignore', invalid='ignore'):
            den = np.log10(vals[..., 0].ravel())
        den = den[np.isfinite(den)]
        density_vmin, density_vmax = np.percentile(den, self.percentile_scaling)
        ui_lo, ui_hi = self._finite_range(den)
        self.update_parameters({'density_vmin': density_vmin, 'density_vmax': density_vmax,
                                'ui_range_density': (ui_lo, ui_hi)})
        self._autorange_using_values(vals[..., 1])
