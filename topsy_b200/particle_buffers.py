"""Device-resident particle data and the per-call draw ranges (reference: src/topsy/particle_buffers.py).

The reference keeps array-of-structures vertex buffers (x,y,z,h | m,q,0 | r,g,b) and writes ``(6, count, 0, first)``
indirect-draw rows for ``multi_draw_indirect``.  Here every attribute is its own contiguous float32 array
(structure-of-arrays: the splat kernel reads four consecutive particles of each array with one 128-bit load) and the
"indirect draw" rows are plain (start, length) ranges handed to ``tsplat_render``.  Physical buffers still follow
``SplitBuffers``.
"""
from __future__ import annotations

import logging

import numpy as np

from . import _native as N
from . import split_buffers

logger = logging.getLogger(__name__)

_UNSET = object()


class ParticleBuffers:
    def __init__(self, loader, device, max_draw_calls_per_buffer: int):
        self.buffers = {}
        self._split_buffers = split_buffers.SplitBuffers(len(loader))
        self._device = device
        self._loader = loader
        self.quantity_name = None
        self._mass_and_quantity_buffers = None
        self._quantity_buffer_is_for_name = _UNSET       # None is a valid name (plain density)
        self._current_vertex_buffers = []
        self._max_draw_calls_per_buffer = max_draw_calls_per_buffer
        self._ranges = [(np.zeros(0, np.int64), np.zeros(0, np.int64)) for _ in range(self._split_buffers.num_buffers)]

    # -- draw ranges ----------------------------------------------------------------------------------------
    def update_particle_ranges(self, particle_mins, particle_lens):
        """Global (start, length) ranges -> per-physical-buffer local ranges (particle_buffers.py:76-82)."""
        per_buf = self._split_buffers.global_to_split_monotonic(particle_mins, particle_lens)
        self._ranges = [(np.asarray(s, dtype=np.int64), np.asarray(l, dtype=np.int64)) for s, l in per_buf]

    def get_ranges(self, bufnum: int):
        return self._ranges[bufnum]

    @property
    def num_buffers(self):
        return self._split_buffers.num_buffers

    # -- attribute arrays -------------------------------------------------------------------------------------
    def _upload_columns(self, columns):
        """columns: list of (N,) host arrays -> per physical buffer a tuple of float32 device tensors."""
        out = []
        for k in range(self._split_buffers.num_buffers):
            a, b = self._split_buffers.buffer_range(k)
            out.append(tuple(self._device.upload(c[a:b]) for c in columns))
        return out

    def _device_columns(self, names):
        """Columns a loader already holds on the device (ArrayDataLoader with the GPU cell layout): split into the
        physical buffers as views, no host round trip.  None when the loader keeps its data on the host."""
        getter = getattr(self._loader, "device_columns", None)
        cols = getter(names) if getter is not None else None
        if cols is None:
            return None
        return self._split_device_columns(cols)

    def _split_device_columns(self, cols):
        out = []
        for k in range(self._split_buffers.num_buffers):
            a, b = self._split_buffers.buffer_range(k)
            out.append(tuple(c[a:b] for c in cols))
        return out

    def get_pos_smooth_buffers(self):
        """Per buffer: (x, y, z, h) float32 tensors (the reference's 'pos_smooth' vertex buffer, :84-91)."""
        if not hasattr(self, "_pos_smooth_buffers"):
            logger.info("Creating position+smoothing buffer")
            self._pos_smooth_buffers = self._device_columns(["x", "y", "z", "h"])
            if self._pos_smooth_buffers is None:
                data = self._loader.get_pos_smooth().astype(np.float32)
                self._pos_smooth_buffers = self._upload_columns([data[:, 0], data[:, 1], data[:, 2], data[:, 3]])
        return self._pos_smooth_buffers

    def get_mass_and_quantity_buffers(self):
        """Per buffer: (m, q) -- q is None for a plain density render (:93-102)."""
        if self._quantity_buffer_is_for_name != self.quantity_name:
            mass_dev = self._device_columns(["m"])
            if mass_dev is not None:
                cols = [self._loader.device_columns(["m"])[0]]
                if self.quantity_name is not None:
                    cols.append(self._loader.device_quantity(self.quantity_name))
                bufs = self._split_device_columns(cols)
            else:
                cols = [np.asarray(self._loader.get_mass(), dtype=np.float32)]
                if self.quantity_name is not None:
                    cols.append(np.asarray(self._loader.get_named_quantity(self.quantity_name), dtype=np.float32))
                bufs = self._upload_columns(cols)
            self._mass_and_quantity_buffers = [b if len(b) == 2 else (b[0], None) for b in bufs]
            self._quantity_buffer_is_for_name = self.quantity_name
        return self._mass_and_quantity_buffers

    def get_rgb_buffers(self):
        """Per buffer: (r, g, b) (:104-111)."""
        if not hasattr(self, "_rgb_masses_buffers"):
            logger.info("Creating rgb buffer")
            self._rgb_masses_buffers = self._device_columns(["rgb_r", "rgb_g", "rgb_b"])
            if self._rgb_masses_buffers is None:
                rgb = np.asarray(self._loader.get_rgb_masses(), dtype=np.float32)
                self._rgb_masses_buffers = self._upload_columns([rgb[:, 0], rgb[:, 1], rgb[:, 2]])
        return self._rgb_masses_buffers

    def specify_vertex_buffer_assignment(self, buffer_names):
        getters = {"pos_smooth": self.get_pos_smooth_buffers, "mass_and_quantity": self.get_mass_and_quantity_buffers,
                   "rgb": self.get_rgb_buffers}
        buffers = []
        for name in buffer_names:
            if name not in getters:
                raise ValueError(f"Unknown buffer name: {name}")
            buffers.append(getters[name]())
        self._current_vertex_buffers = buffers

    def issue_draw(self, engine, mode: int, clear: bool, image=None):
        """Replaces issue_draw_indirect (:70-74): one tsplat_render per physical buffer with the current ranges."""
        pos_bufs, weight_bufs = self._current_vertex_buffers
        for k in range(self._split_buffers.num_buffers):
            starts, lens = self._ranges[k]
            if len(starts) == 0 and not clear:
                continue
            engine.set_particles(*pos_bufs[k])
            weights = weight_bufs[k]
            if mode == N.MODE_SURFACE and weights[1] is None:
                # no quantity selected: the reference's (m, q, 0) vertex buffer carries q = 0
                if not hasattr(self, "_zero_quantity"):
                    self._zero_quantity = {}
                if k not in self._zero_quantity:
                    import torch
                    self._zero_quantity[k] = torch.zeros_like(weights[0])
                weights = (weights[0], self._zero_quantity[k])
            engine.set_weights(*weights)
            engine.render(mode, starts, lens, clear=clear, image=image)
            clear = False
