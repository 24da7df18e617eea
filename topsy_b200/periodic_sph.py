"""Periodic tiling of the rendered box (reference: src/topsy/periodic_sph.py).

The snapshot is splatted once; the presented image is the sum of copies of that image displaced by every lattice vector of
the periodic box that lands within one box length in depth, faded out between half a box and one box away so that
rotating the view does not pop replicas in and out.  The reference draws instanced textured quads with additive blending
(overlay.wgsl); here one gather kernel (``tsplat_periodic_accumulate``) sums the bilinearly sampled replicas per pixel.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _native as N
from . import sph
from .device import Texture
from .drawreason import DrawReason


def replica_offsets_and_weights(rotation_matrix, panel_scale, num_repetitions=2):
    """Clip-space (x, y) shifts and weights of the box replicas to draw (periodic_sph.py:36-54)."""
    n = num_repetitions
    grid = np.stack(np.meshgrid(np.arange(-n, n + 1), np.arange(-n, n + 1), np.arange(-n, n + 1), indexing="ij"),
                    axis=-1).reshape(-1, 3).astype(np.float32)
    rotated = grid @ np.asarray(rotation_matrix, dtype=np.float32).T
    depth = np.abs(rotated[:, 2])
    keep = depth < 1.0
    weights = np.where(depth > 0.5, 1.0 - 2.0 * (depth - 0.5), 1.0).astype(np.float32)
    return (rotated[keep, :2] * np.float32(panel_scale)).astype(np.float32), weights[keep]


class PeriodicSPH(sph.SPH):
    def __init__(self, visualizer, render_size):
        super().__init__(visualizer, render_size, wrapping=True)
        self._periodic_images = {}
        self._periodic_texture = Texture(self._current_periodic_image, self.render_format, "proxy_sph")
        self.num_repetitions = 2

    def _current_periodic_image(self):
        src = self._current_image()
        key = src.shape[2]
        if key not in self._periodic_images:
            self._periodic_images[key] = torch.zeros_like(src)
        return self._periodic_images[key]

    def get_output_texture(self) -> Texture:
        return self._periodic_texture

    def render(self, draw_reason=DrawReason.CHANGE):
        if draw_reason == DrawReason.PRESENTATION_CHANGE:
            return
        super().render(draw_reason)
        panel_scale = self._visualizer.periodicity_scale / self._visualizer.scale
        offsets, weights = replica_offsets_and_weights(self.rotation_matrix, panel_scale, self.num_repetitions)
        src = self._current_image()
        dst = self._current_periodic_image()
        eng = self._engine
        offsets = np.ascontiguousarray(offsets, np.float32); weights = np.ascontiguousarray(weights, np.float32)
        stream = ctypes.c_void_p(torch.cuda.current_stream(self._device.torch_device).cuda_stream)
        N.check(eng.lib.tsplat_periodic_accumulate(eng._ctx, ctypes.c_void_p(src.data_ptr()), ctypes.c_void_p(dst.data_ptr()),
                                                   src.shape[2], offsets.ctypes.data_as(ctypes.c_void_p),
                                                   weights.ctypes.data_as(ctypes.c_void_p), len(weights), stream))

    def _get_image_unscaled(self):
        if not self.has_rendered:
            self.render(DrawReason.EXPORT)
        img = self._current_periodic_image().cpu().numpy()
        if img.shape[2] < self._nchannels_output:
            padded = np.zeros(img.shape[:2] + (self._nchannels_output,), dtype=self._output_dtype)
            padded[..., :img.shape[2]] = img
            img = padded
        return img
