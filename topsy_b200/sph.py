"""SPH projection renderers (reference: src/topsy/sph.py + src/topsy/shaders/sph.wgsl).

Same classes, constructor, attributes and methods as the reference so that ``Visualizer`` and user code can switch
without edits; the wgpu render pass (instanced quads + additive blending) is replaced by the hand-written CUDA kernels
behind ``tsplat_render`` (topsy_b200/csrc/tsplat.cu).

  SPH / BivariateSPH   2-channel accumulation (K m/h^2, K m/h^2 q); when no quantity is selected only the density
                       channel is accumulated (1-channel image, 20 B/particle instead of 24) and the second channel
                       of ``get_image()`` is returned as zeros, as in the reference
  RGBSPH               (K r/h^2, K g/h^2, K b/h^2, fragment count)
  DepthSPH             (K m/h^2, K m/h^2 z_clip)  -> get_depth_image()
  DepthSPHWithOcclusion  z-buffered (quantity, depth) of the front-most particles above a density cut ('surface' mode)
"""
from __future__ import annotations

import copy
import logging

import numpy as np

from . import _native as N
from . import config, performance
from .camera import transform_matrix
from .device import Texture
from .drawreason import DrawReason
from .util import TimeGpuOperation

logger = logging.getLogger(__name__)


class SPH:
    render_format = "rg32float"
    _nchannels_input = 2
    _nchannels_output = 2
    _output_dtype = np.float32
    _buffer_name = "mass_and_quantity"

    # uniform block of the reference's vertex shader, kept so that ``last_transform_params`` has the same fields
    _transform_params_dtype = [("transform", np.float32, (4, 4)),
                               ("scale_factor", np.float32, (1,)),
                               ("min_max_size", np.float32, (2,)),
                               ("boxsize_by_2_clipspace", np.float32, (1,)),
                               ("density_cut", np.float32, (1,))]

    def __init__(self, visualizer, render_resolution, wrapping=False, share_render_progression=None, exchange_cache=None):
        logger.info(f"Initializing {self.__class__} with resolution {render_resolution}")
        self._visualizer = visualizer
        self._render_resolution = render_resolution
        self._device = visualizer.device
        self._wrapping = wrapping
        self._engine = self._device.engine(render_resolution)
        self._render_texture = Texture(self._current_image, self.render_format, "sph_render_texture")
        self._render_timer = TimeGpuOperation(self._device)
        if share_render_progression is not None:
            self._render_progression = share_render_progression
        else:
            self._render_progression = visualizer.data_loader.get_render_progression()
        self._images = {}
        self._last_mode = None
        # multi-GPU (no reference counterpart: topsy is single-device): under torchrun every rank holds a stripe of the
        # particles (distributed.shard_loader); the renderer splats into a partial image and all-reduces it over NVLink
        # after every frame, so that everything downstream (get_image, colormap, autorange) sees the full image.
        from . import distributed
        self._shard_rank, self._shard_world = distributed.shard_context()
        self._exchanges = {} if exchange_cache is None else exchange_cache

        self.scale = config.DEFAULT_SCALE
        self.min_pixels = 0.0       # kept for API compatibility: like in the reference these have no effect
        self.max_pixels = np.inf
        self.rotation_matrix = np.eye(3)
        self.position_offset = np.zeros(3)
        self.has_rendered = False

    # -- mode / image management -------------------------------------------------------------------------------
    def _mode(self) -> int:
        has_quantity = self._visualizer.particle_buffers.quantity_name is not None
        return N.MODE_WEIGHTED if has_quantity else N.MODE_DENSITY

    def _image_for_mode(self, mode):
        """Each renderer owns its accumulation images (one per channel count) so that a throw-away DepthSPH or a second
        Visualizer at the same resolution never clobbers this one's progressive state."""
        import torch
        channels = N.MODE_CHANNELS[mode]
        if self._shard_world > 1:
            return self._exchange_for(channels).partial
        if channels not in self._images:
            self._images[channels] = torch.zeros((self._render_resolution, self._render_resolution, channels),
                                                 dtype=torch.float32, device=self._device.torch_device)
        return self._images[channels]

    def _exchange_for(self, channels):
        if channels not in self._exchanges:
            from . import distributed
            self._exchanges[channels] = distributed.ImageExchange(self._engine, self._render_resolution, channels)
        return self._exchanges[channels]

    def _current_image(self):
        mode = self._last_mode if self._last_mode is not None else self._mode()
        if self._shard_world > 1:
            return self._exchange_for(N.MODE_CHANNELS[mode]).reduced      # the all-reduced image, identical on every rank
        return self._image_for_mode(mode)

    def get_output_texture(self) -> Texture:
        return self._render_texture

    # -- camera ---------------------------------------------------------------------------------------------
    def _get_transform_params(self):
        """Structured record with the reference's uniform layout (sph.py:268-299); ``transform`` is stored transposed
        (column-major) exactly as the reference uploads it."""
        M = transform_matrix(self.rotation_matrix, self.position_offset, self.scale)
        params = np.zeros((), dtype=self._transform_params_dtype)
        params["transform"] = M.T
        params["scale_factor"] = 1.0 / self.scale
        period = self._visualizer.periodicity_scale
        params["boxsize_by_2_clipspace"] = 0.5 * period / self.scale if period is not None else 0.0
        res = self._render_resolution
        params["min_max_size"] = (2.0 * self.min_pixels / res, 2.0 * self.max_pixels / res)
        return params

    def _update_transform_buffer(self):
        params = self._get_transform_params()
        self.last_transform_params = params
        self._engine.set_camera(np.ascontiguousarray(params["transform"].T), float(params["scale_factor"][0]))

    def _reassert_engine_state(self):
        """The engine is shared by every renderer at this resolution: re-assert our camera before adding blocks."""
        self._engine.set_camera(np.ascontiguousarray(self.last_transform_params["transform"].T),
                                float(self.last_transform_params["scale_factor"][0]))

    # -- rendering ------------------------------------------------------------------------------------------
    def invalidate(self, draw_reason=DrawReason.CHANGE):
        if draw_reason not in (DrawReason.REFINE, DrawReason.PRESENTATION_CHANGE):
            self.has_rendered = False

    def render(self, draw_reason=DrawReason.CHANGE):
        """One frame of the progressive render: a sequence of blocks chosen by the progression (sph.py:306-332)."""
        performance.signposter.emit_event("Start SPH render")
        if draw_reason == DrawReason.PRESENTATION_CHANGE:
            return
        mode = self._mode()
        if draw_reason != DrawReason.REFINE or mode != self._last_mode:
            self._render_progression.select_sphere(-self.position_offset, self.scale * 1.2)
            self._update_transform_buffer()
        else:
            self._reassert_engine_state()
        self._last_mode = mode
        image = self._image_for_mode(mode)
        buffers = self._visualizer.particle_buffers
        buffers.specify_vertex_buffer_assignment(['pos_smooth', self._buffer_name])

        clear = self._render_progression.start_frame(draw_reason)
        # an EXPORT frame's block sizes do not depend on the elapsed time (progressive_render.py:67-70): its blocks are
        # enqueued back to back and the host waits once, when the frame's time is handed to the progression below
        wait = draw_reason != DrawReason.EXPORT
        while block := self._render_progression.get_block(self._render_timer.total_time_in_frame(wait)):
            buffers.update_particle_ranges(*block)
            with self._render_timer:
                buffers.issue_draw(self._engine, mode, clear, image=image)
            self._render_progression.end_block(self._render_timer.total_time_in_frame(wait))
            clear = False
        if not wait:
            self._render_progression.set_time_in_frame(self._render_timer.total_time_in_frame())
        self._render_timer.end_frame()

        self.last_render_mass_scale = self._render_progression.end_frame_get_scalefactor()
        if self._shard_world > 1:
            # reduced = sum over ranks of (rank's mass scale) * (rank's partial image): already an estimate of the full
            # image, so nothing is left to rescale afterwards
            self._exchange_for(N.MODE_CHANNELS[mode]).allreduce(self.last_render_mass_scale, zmax=(mode == N.MODE_SURFACE))
            self.last_render_mass_scale = 1.0
        self.last_render_fps = 1.0 / max(self._render_timer.running_mean_duration, 1e-9)
        self.has_rendered = True

    def needs_refine(self):
        local = self._render_progression.needs_refine()
        if self._shard_world > 1:
            # frames are collective (every render ends in the image all-reduce): all ranks refine until the last is done
            import torch
            import torch.distributed as dist
            flag = torch.tensor([1 if local else 0], dtype=torch.int32, device=self._device.torch_device)
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)
            return bool(int(flag.item()))
        return local

    # -- readback -------------------------------------------------------------------------------------------
    def _get_image_unscaled(self):
        if not self.has_rendered:
            logger.info("Export-quality render has been triggered, because no render has been done yet.")
            self.render(DrawReason.EXPORT)
        img = self._current_image().cpu().numpy()
        if img.shape[2] < self._nchannels_output:          # density-only accumulation: quantity channel is all zero
            padded = np.zeros(img.shape[:2] + (self._nchannels_output,), dtype=self._output_dtype)
            padded[..., :img.shape[2]] = img
            img = padded
        return img

    def get_image(self) -> np.ndarray:
        """Last rendered image, (R, R, channels) float32, rescaled to the full particle count (sph.py:118-125)."""
        return self._get_image_unscaled() * self.last_render_mass_scale

    def _get_depth_renderer(self):
        if not hasattr(self, "_depth_exchanges"):
            self._depth_exchanges = {}          # the throw-away depth renderers share one set of symmetric images
        renderer = DepthSPH(self._visualizer, self._render_resolution, wrapping=self._wrapping,
                            share_render_progression=copy.copy(self._render_progression),
                            exchange_cache=self._depth_exchanges)
        renderer.rotation_matrix = self.rotation_matrix
        renderer.position_offset = self.position_offset
        renderer.scale = self.scale
        return renderer

    def get_depth_image(self, depth_renderer_reason=DrawReason.CHANGE) -> np.ndarray:
        """Mass-weighted mean depth of the scene in simulation units, used to pick a point under the cursor
        (sph.py:97-116).  CHANGE renders a quick subsample; EXPORT every particle."""
        depth_renderer = self._get_depth_renderer()
        depth_renderer.render(depth_renderer_reason)
        image = depth_renderer.get_image()
        with np.errstate(divide='ignore', invalid='ignore'):
            depth_viewport = image[..., 1] / image[..., 0]
        return (depth_viewport - 0.5) * self.scale * 2.0


class BivariateSPH(SPH):
    """Renders a (density, mass-weighted mean) pair -- same accumulation as SPH."""


class RGBSPH(SPH):
    render_format = "rgba32float"
    _buffer_name = 'rgb'
    _nchannels_input = 3
    _nchannels_output = 4

    def _mode(self):
        return N.MODE_RGB


class DepthSPH(SPH):
    """Second channel accumulates K m/h^2 * z_clip (sph.wgsl:85-91)."""

    def _mode(self):
        return N.MODE_DEPTH


class LocalSphereKernel:
    """Depth of a sphere of radius 2h below its silhouette, -0.01 outside it (sph.py:446-455); tabulated for the device by
    ``kernel_lut.local_sphere_lut``."""

    def get_value(self, distance):
        return np.sqrt(4.0 - distance ** 2) if distance < 2.0 else -0.01


class DepthSPHWithOcclusion(SPH):
    """Renders the front-most particles above a density cut: per pixel (quantity, depth) of the fragment nearest to the
    camera (reference: sph.py:457-601; vertex_depth_with_cut / fragment_raw, sph.wgsl:93-158).  The reference's depth
    attachment + depth_compare=greater become one 64-bit atomic max per fragment on the (quantity, depth) pixel itself
    (kernel K9, topsy_b200/csrc/tsplat_surface.cuh)."""
    _nchannels_output = 2
    _rho_percentiles_num_samples = 101      # the density cut is tabulated at every percentile from 0 to 100

    def __init__(self, visualizer, render_resolution, wrapping=False, share_render_progression=None, exchange_cache=None):
        super().__init__(visualizer, render_resolution, wrapping, share_render_progression, exchange_cache)
        mass = self._visualizer.data_loader.get_mass()
        smooth = self._visualizer.data_loader.get_smooth()
        rho = mass / smooth ** 3
        self._cut_min = np.log10(rho.min())
        self._cut_max = np.log10(rho.max())
        self._percentile_to_den_cut = np.quantile(rho, np.linspace(0, 1, self._rho_percentiles_num_samples))
        self._cut_val = np.mean(self.get_density_cut_percentile_range())     # start at the median density
        from .kernel_lut import local_sphere_lut
        self._engine.set_surface(local_sphere_lut(), 0.0)

    def _mode(self):
        return N.MODE_SURFACE

    def _get_transform_params(self):
        tp = super()._get_transform_params()
        tp["density_cut"] = self._percentile_to_den_cut[int(self._cut_val / 100.0 * (self._rho_percentiles_num_samples - 1))]
        return tp

    def _update_transform_buffer(self):
        super()._update_transform_buffer()
        self._engine.set_surface(None, float(self.last_transform_params["density_cut"][0]))

    def _reassert_engine_state(self):
        super()._reassert_engine_state()
        self._engine.set_surface(None, float(self.last_transform_params["density_cut"][0]))

    def get_density_cut_percentile(self):
        return self._cut_val

    def set_density_cut_percentile(self, value):
        self._cut_val = value

    def get_density_cut_percentile_range(self):
        return 0.0, 100.0

    def get_image(self):
        return self._get_image_unscaled()       # maxima, not sums: no rescaling for partial renders
