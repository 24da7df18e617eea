"""Splitting a logical particle array over several physical buffers (reference: src/topsy/split_buffers.py).

topsy splits because wgpu caps the size of one buffer; here a "physical buffer" is one PyTorch CUDA allocation, and in
the multi-GPU layout (topsy_b200/distributed.py) buffer k lives on GPU k.  The address arithmetic -- 'global' particle
index <-> (buffer number, offset) -- is the same contract, including the exceptions.
"""
from __future__ import annotations

import logging

import numpy as np

from . import config, performance

logger = logging.getLogger(__name__)


class SplitBuffers:
    def __init__(self, num_particles: int, max_particles_per_buffer: int | None = None):
        if max_particles_per_buffer is None:
            max_particles_per_buffer = config.MAX_PARTICLES_PER_BUFFER
        self._num_particles = num_particles
        self._max_particles_per_buffer = max_particles_per_buffer
        n_buf = max(1, -(-num_particles // max_particles_per_buffer))
        self._num_buffers = n_buf
        sizes = np.full(n_buf, max_particles_per_buffer, dtype=np.intp)
        sizes[-1] = num_particles - (n_buf - 1) * max_particles_per_buffer
        self._buffer_particle_sizes = sizes
        self._buffer_particle_starts = np.cumsum(sizes) - sizes
        logger.info(f"Splitting {num_particles} particles into {n_buf} buffer(s)")

    @property
    def num_buffers(self) -> int:
        return self._num_buffers

    def _global_to_split_address(self, address: int):
        buf = np.searchsorted(self._buffer_particle_starts, address, side='right') - 1
        return buf, address - self._buffer_particle_starts[buf]

    def global_to_split(self, start: int, length: int):
        """(buffer numbers, local starts, lengths) covering the global range; ValueError if it runs off the end."""
        bufs, starts, lengths = [], [], []
        buf, local = self._global_to_split_address(start)
        left = length
        while left > 0 and buf < self._num_buffers:
            take = min(left, self._buffer_particle_sizes[buf] - local)
            bufs.append(buf); starts.append(local); lengths.append(take)
            left -= take
            buf += 1
            local = 0
        if left > 0:
            raise ValueError(f"Requested length {length} starting at {start} exceeds available buffers")
        return bufs, starts, lengths

    def global_to_split_monotonic(self, start, length):
        """Per-buffer ``(starts, lengths)`` lists for monotonically increasing global ranges (split_buffers.py:78-116).
        Always returns ``num_buffers`` entries.  The reference sweeps the ranges in a Python loop; this runs once per
        render block with up to n_cells ranges, so it is vectorised: every buffer clips all ranges at once."""
        performance.signposter.emit_event("global_to_split_monotonic")
        start = np.asarray(start, dtype=np.int64).ravel()
        end = start + np.asarray(length, dtype=np.int64).ravel()
        live = end > start
        if live.any() and end[live].max() > self._num_particles:
            bad = int(np.argmax(live & (end > self._num_particles)))
            raise ValueError(f"Requested length {int(end[bad] - start[bad])} starting at {int(start[bad])} exceeds available buffers")
        result = []
        for k in range(self._num_buffers):
            b0 = int(self._buffer_particle_starts[k])
            b1 = b0 + int(self._buffer_particle_sizes[k])
            lo = np.maximum(start, b0)
            n = np.minimum(end, b1) - lo
            keep = n > 0
            result.append(((lo[keep] - b0).tolist(), n[keep].tolist()))
        performance.signposter.emit_event("end global_to_split_monotonic")
        return result

    # -- storage --------------------------------------------------------------------------------------------
    def create_buffers(self, device, item_size: int, usage=None):
        """One uint8 CUDA tensor of ``size * item_size`` bytes per physical buffer (``usage`` is accepted for
        signature compatibility with the wgpu version and ignored)."""
        return [device.create_buffer(int(n) * item_size) for n in self._buffer_particle_sizes]

    def write_buffers(self, device, buffers, data: np.ndarray) -> None:
        if len(buffers) != self._num_buffers:
            raise ValueError(f"Number of buffers {len(buffers)} does not match number of split buffers {self._num_buffers}")
        if len(data) != self._num_particles:
            raise ValueError(f"Data size {len(data)} does not match number of particles {self._num_particles}")
        for k, buf in enumerate(buffers):
            first = self._buffer_particle_starts[k]
            device.write_buffer(buf, data[first:first + self._buffer_particle_sizes[k]])

    def buffer_range(self, k: int):
        first = int(self._buffer_particle_starts[k])
        return first, first + int(self._buffer_particle_sizes[k])
