"""Orchestration: owns the device, the data, the SPH renderer and the colormap (reference: src/topsy/visualizer.py).

Constructor, properties and methods follow the reference so scripts and tests written against topsy run unchanged on
the hot path: ``render_mode``, ``scale`` / ``rotation_matrix`` / ``position_offset`` / ``quantity_name``, ``rotate``,
``draw``, ``render_sph``, ``get_sph_image``, ``get_sph_presentation_image``, ``get_presentation_image``,
``get_depth_image``, ``colormap_autorange``, ``save``.  Decorations that the reference builds with matplotlib
(colorbar, scalebar, status text, crosshair lines) are outside the hot path: they are represented by light records so
that the attributes exist, but nothing is drawn for them.
"""
from __future__ import annotations

import logging
import time
from contextlib import contextmanager
from dataclasses import dataclass

import numpy as np

from . import canvas as canvas_module
from . import colormap, config, loader, particle_buffers, sph
from .camera import rotation_about_x, rotation_about_y
from .device import Device, Texture
from .drawreason import DrawReason

logger = logging.getLogger(__name__)


@dataclass
class ColorbarInfo:
    """What the reference's ColorbarOverlay displays (colorbar.py); kept as data only."""
    vmin: float
    vmax: float
    colormap_name: str
    label: str


class VisualizerBase:
    colorbar_aspect_ratio = config.COLORBAR_ASPECT_RATIO
    show_status = True
    device = None        # shared by all instances, like the reference's wgpu device

    def __init__(self, data_loader_class=loader.TestDataLoader, data_loader_args=(), data_loader_kwargs={},
                 *, render_resolution=config.DEFAULT_RESOLUTION, periodic_tiling=False,
                 colormap_name=config.DEFAULT_COLORMAP, canvas_class=None, render_mode='univariate'):
        self._render_resolution = render_resolution
        self._colorbar = None
        self._sph = None
        self._colormap = None
        self.crosshairs_visible = False
        self._prevent_sph_rendering = False
        self.show_colorbar = True
        self.show_scalebar = True
        self._validate_render_mode(render_mode)
        self._render_mode = render_mode
        if canvas_class is None:
            canvas_class = canvas_module.VisualizerCanvas
        self.canvas = canvas_class(visualizer=self, title="topsy")
        self._setup_device()
        self._configure_canvas_context()
        self._initialize_data_loader_and_buffers(data_loader_class, data_loader_args, data_loader_kwargs)
        self._periodic_tiling = periodic_tiling
        self._status_text = "topsy"
        self._initialize_sph_and_colormap_and_bar(colormap_name)
        self._last_status_update = 0.0

    # -- construction helpers -----------------------------------------------------------------------------------
    def _setup_device(self):
        if type(self).device is None:
            VisualizerBase.device = Device()
        self.context = self.canvas.get_context("wgpu")

    def _render_mode_to_canvas_format(self, render_mode):
        if render_mode is None:
            return None
        if render_mode.endswith('hdr'):
            return "rgba16float"
        fmt = self.context.get_preferred_format(None)
        return fmt[:-5] if fmt.endswith("-srgb") else fmt

    def _configure_canvas_context(self):
        self.canvas_format = self._render_mode_to_canvas_format(self._render_mode)
        self.context.configure(device=self.device, format=self.canvas_format)
        logger.info(f"Canvas format {self.canvas_format}")

    def _initialize_data_loader_and_buffers(self, data_loader_class, data_loader_args, data_loader_kwargs):
        self.data_loader = data_loader_class(self.device, *data_loader_args, **data_loader_kwargs)
        # multi-GPU: under torchrun every rank keeps its per-cell stripe of the snapshot (no reference counterpart; the
        # wiring extended here is visualizer.py:75-80 loader -> ParticleBuffers)
        from . import distributed
        rank, world = distributed.shard_context()
        distributed.shard_loader(self.data_loader, rank, world)
        regions = self.data_loader.get_render_progression().get_max_particle_regions_per_block()
        self.particle_buffers = particle_buffers.ParticleBuffers(self.data_loader, self.device, regions)
        self.periodicity_scale = self.data_loader.get_periodicity_scale()

    def _get_sph_class_for_render_mode(self, render_mode):
        if render_mode in ('rgb', 'rgb-hdr'):
            return sph.RGBSPH
        if render_mode == 'surface':
            return sph.DepthSPHWithOcclusion
        return sph.SPH

    def _get_colormap_parameters_for_render_mode(self, render_mode):
        params = {'weighted_average': self.quantity_name is not None}
        if render_mode == 'rgb':
            params.update({'type': 'rgb', 'hdr': False, 'log': True})
        elif render_mode == 'rgb-hdr':
            params.update({'type': 'rgb', 'hdr': True, 'log': True})
        elif render_mode == 'bivariate':
            params.update({'type': 'bivariate'})
        elif render_mode == 'surface':
            params.update({'type': 'surface'})
        else:
            params.update({'type': 'density'})
        return params

    def _initialize_sph_and_colormap_and_bar(self, colormap_name=None):
        """(Re-)create renderer, colormap and colorbar, keeping the camera (visualizer.py:122-154)."""
        old = (self._sph.rotation_matrix, self._sph.position_offset, self._sph.scale) if self._sph is not None \
            else (None, None, None)
        if self._periodic_tiling:
            from . import periodic_sph
            self._sph = periodic_sph.PeriodicSPH(self, self._render_resolution)
        else:
            sph_class = self._get_sph_class_for_render_mode(self._render_mode)
            logger.info(f"Using {sph_class.__name__} renderer for render mode '{self._render_mode}'")
            self._sph = sph_class(self, self._render_resolution)
        self.reset_view(rotation_matrix=old[0], position_offset=old[1], scale=old[2])
        self.invalidate()
        if colormap_name is None:
            colormap_name = self._colormap.get_parameter('colormap_name')
        self.render_texture = self._sph.get_output_texture()
        self._colormap = colormap.ColormapHolder(self.device, self.render_texture, self.canvas_format)
        self._colormap.update_parameters({'colormap_name': colormap_name})
        self._initialize_colormap_and_bar()

    def _initialize_colormap_and_bar(self):
        params = self._get_colormap_parameters_for_render_mode(self._render_mode)
        changed_type = self._colormap.update_parameters(params)
        params = self._colormap.get_parameters()
        show_colorbar = (params['type'] not in ('rgb', 'surface')
                         or (params['type'] == 'surface' and params['weighted_average']))
        if changed_type or params['vmin'] is None or params['vmax'] is None:
            logger.info("Autorange colormap parameters")
            self._autorange()
            params = self._colormap.get_parameters()
        if show_colorbar:
            self._colorbar = ColorbarInfo(params['vmin'], params['vmax'], params['colormap_name'], self._get_colorbar_label())
        else:
            self._colorbar = None

    def _get_colorbar_label(self):
        label = self.data_loader.get_quantity_label(self.quantity_name)
        if self._colormap.get_parameter('log'):
            label = r"$\log_{10}$ " + label
        return label

    # -- invalidation / camera ------------------------------------------------------------------------------------
    def invalidate(self, reason=DrawReason.CHANGE):
        self._sph.invalidate(reason)
        self.canvas.request_draw(lambda: self.draw(reason))

    def rotate(self, x_angle, y_angle):
        self.rotation_matrix = rotation_about_y(x_angle) @ rotation_about_x(y_angle) @ self.rotation_matrix

    _x_rotation_matrix = staticmethod(rotation_about_y)      # the reference's (misleading) names, visualizer.py:347-357
    _y_rotation_matrix = staticmethod(rotation_about_x)

    @property
    def colormap(self):
        return self._colormap

    @property
    def rotation_matrix(self):
        return self._sph.rotation_matrix

    @rotation_matrix.setter
    def rotation_matrix(self, value):
        self._sph.rotation_matrix = value
        self.invalidate()

    @property
    def position_offset(self):
        return self._sph.position_offset

    @position_offset.setter
    def position_offset(self, value):
        self._sph.position_offset = value
        self.invalidate()

    @property
    def scale(self):
        """Half-width of the view in simulation length units."""
        return self._sph.scale

    @scale.setter
    def scale(self, value):
        self._sph.scale = value
        self.invalidate()

    def reset_view(self, rotation_matrix=None, position_offset=None, scale=None):
        if rotation_matrix is None:
            rotation_matrix = np.eye(3)
        if position_offset is None:
            position_offset = -self.data_loader.get_initial_center()
        if scale is None:
            scale = self.data_loader.get_initial_view_width()
        self._sph.rotation_matrix = rotation_matrix
        self._sph.scale = scale
        self._sph.position_offset = position_offset

    # -- render mode ------------------------------------------------------------------------------------------
    def _validate_render_mode(self, new_render_mode):
        valid_modes = {'univariate', 'bivariate', 'rgb', 'rgb-hdr', 'surface'}
        if new_render_mode not in valid_modes:
            raise ValueError(f"Invalid render_mode '{new_render_mode}'. Valid modes: {valid_modes}")

    @property
    def render_mode(self):
        return self._render_mode

    @render_mode.setter
    def render_mode(self, value):
        self._update_render_mode(value)

    def _update_render_mode(self, new_render_mode, revert_on_failure=True):
        self._validate_render_mode(new_render_mode)
        old_render_mode = getattr(self, "_render_mode", None)
        self._render_mode = new_render_mode
        try:
            if self._render_mode_to_canvas_format(old_render_mode) != self._render_mode_to_canvas_format(new_render_mode):
                self._configure_canvas_context()
            self._initialize_sph_and_colormap_and_bar()
        except Exception:
            if revert_on_failure:
                logger.error(f"Failed to update render mode to '{new_render_mode}'; reverting to '{old_render_mode}'")
                self._update_render_mode(old_render_mode, revert_on_failure=False)
            raise
        self.invalidate(DrawReason.CHANGE)

    # -- quantity ---------------------------------------------------------------------------------------------
    @property
    def quantity_name(self):
        """Name of the quantity shown as a mass-weighted mean, or None for projected density."""
        return self.particle_buffers.quantity_name

    @property
    def averaging(self):
        return self.quantity_name is not None

    @quantity_name.setter
    def quantity_name(self, value):
        if value == self.particle_buffers.quantity_name:
            return
        if value is not None:
            try:
                self.data_loader.get_named_quantity(value)
            except Exception as e:
                raise ValueError(f"Unable to get quantity named '{value}'") from e
        self.particle_buffers.quantity_name = value
        self.invalidate(DrawReason.CHANGE)
        self._colormap.update_parameters({'vmin': None, 'vmax': None, 'log': None})
        self._initialize_colormap_and_bar()

    use_device_autorange = True      # False: read the image back and use numpy percentiles, like the reference

    def _autorange(self):
        if self.use_device_autorange:
            if not self._sph.has_rendered:
                self._sph.render(DrawReason.EXPORT)
            self._colormap.autorange_texture(self._sph.last_render_mass_scale)
        else:
            self._colormap.autorange(self._sph.get_image())

    def colormap_autorange(self):
        self._autorange()
        self.invalidate(DrawReason.PRESENTATION_CHANGE)

    # -- drawing ----------------------------------------------------------------------------------------------
    @contextmanager
    def prevent_sph_rendering(self):
        self._prevent_sph_rendering = True
        try:
            yield
        finally:
            self._prevent_sph_rendering = False

    def draw(self, reason, target_texture_view=None):
        """One presented frame: (progressive) SPH render, then the colormap pass into the target (visualizer.py:386-402)."""
        if target_texture_view is None:
            target_texture_view = self.canvas.get_context("wgpu").get_current_texture().create_view()
        if not self._prevent_sph_rendering:
            self.render_sph(reason)
        self._colormap.set_scaling(*target_texture_view.size[:2], self._sph.last_render_mass_scale)
        self._colormap.encode_render_pass(None, target_texture_view)
        self._update_status()
        if reason != DrawReason.EXPORT and not self._prevent_sph_rendering and self._sph.needs_refine():
            self.invalidate(DrawReason.REFINE)

    def render_sph(self, draw_reason=DrawReason.CHANGE):
        self._sph.render(draw_reason)

    def display_status(self, text, timeout=0.5):
        self._override_status_text = text
        self._override_status_text_until = time.time() + timeout

    def _update_status(self):
        """Status line content of the reference ("N fps /X.Xds /Y.Ygf", visualizer.py:438-448), kept as a string."""
        if hasattr(self._sph, 'last_render_fps'):
            text = f"${self._sph.last_render_fps:.0f}$ fps"
            factor = np.round(self._sph.last_render_mass_scale, 1)
            if factor > 1.1:
                text += f" /{factor:.1f}ds"
            geom = self._sph._render_progression.get_fraction_volume_selected()
            if geom < 0.9:
                text += f" /{1. / geom:.1f}gf"
            self._status_text = text

    # -- export -----------------------------------------------------------------------------------------------
    def get_sph_image(self) -> np.ndarray:
        """Logical content of the SPH image: density (R,R), weighted mean (R,R), bivariate (R,R,2) or rgb (R,R,3)."""
        return self._colormap.sph_raw_output_to_content(self._sph.get_image())

    def get_sph_presentation_image(self) -> np.ndarray:
        """EXPORT-quality render + colormap at the render resolution: (R,R,4) uint8, or float16 in 'rgb-hdr'."""
        res = self._render_resolution
        texture = self.device.create_texture((res, res, 1), self.canvas_format, label="output_texture")
        self.render_sph(DrawReason.EXPORT)
        self._colormap.set_scaling(res, res, self._sph.last_render_mass_scale)
        self._colormap.encode_render_pass(None, texture.create_view())
        return self._texture_to_rgba_numpy(texture)

    def get_depth_image(self) -> np.ndarray:
        return self._sph.get_depth_image()

    def sph_clipspace_to_screen_clipspace_matrix(self):
        """Scaling that fits the square SPH image over the larger window dimension (visualizer.py:407-422); the colormap
        and surface kernels apply the same mapping through ``window_aspect_ratio``."""
        aspect_ratio = self.canvas.width_physical / self.canvas.height_physical
        matr = np.eye(4, dtype=np.float32)
        matr[0, 0] = 1.0 if aspect_ratio >= 1 else 1.0 / aspect_ratio
        matr[1, 1] = aspect_ratio if aspect_ratio > 1 else 1.0
        return matr

    def show(self, force=False):
        """The reference opens the Qt / Jupyter window here (visualizer.py:572-591).  Windowing is outside the B200 hot path:
        the offscreen canvas is returned so that ``vis.show()`` in a script keeps working."""
        return self.canvas

    def get_presentation_image(self, resolution=(640, 480)) -> np.ndarray:
        """What the window would show at ``resolution`` (without the matplotlib decorations of the reference)."""
        texture = self.device.create_texture((resolution[0], resolution[1], 1), self.canvas_format, label="output_texture")
        self.draw(DrawReason.EXPORT, texture.create_view())
        return self._texture_to_rgba_numpy(texture)

    def _texture_to_rgba_numpy(self, texture: Texture):
        if not (texture.format.endswith("8unorm") or texture.format.endswith("16float")):
            raise ValueError(f"Unsupported texture format {texture.format}")
        result = texture.tensor.cpu().numpy()
        if texture.format.startswith("bgr"):
            result = result[..., [2, 1, 0, 3]]
        return result

    def save(self, filename='output.npy'):
        """``.npy`` -> logical SPH content; any image suffix PIL knows -> the colormapped render (visualizer.py:528-570;
        the reference decorates image files with matplotlib axes, which are outside the hot path)."""
        self._sph.render(DrawReason.EXPORT)
        if filename.endswith(".npy"):
            np.save(filename, self.get_sph_image())
            return
        image = self.get_sph_presentation_image()
        if image.dtype != np.uint8:
            image = (np.clip(image.astype(np.float32), 0, 1) * 255).astype(np.uint8)
        from PIL import Image
        Image.fromarray(image, mode="RGBA").save(filename)


class Visualizer(VisualizerBase):
    pass
