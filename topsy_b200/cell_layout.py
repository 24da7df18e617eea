"""Spatial bucketing of a snapshot into nside^3 cubic cells (reference: src/topsy/cell_layout.py).

The layout is what lets the renderer (a) pick only the cells near the camera (``cells_in_sphere``) and (b) take the
same *fraction* of every cell for a progressive frame.  ``from_positions`` reproduces the reference's cell assignment
bit-for-bit -- the index arithmetic runs in the dtype of the positions with the same operation order -- either with
numpy on the host or, for device-resident positions, with the K4 counting-sort kernels behind ``tsplat_cell_layout``.
numpy's default argsort is not stable, so the order of particles *within* a cell is unspecified by the reference; both
paths here return the stable order (ascending original index inside each cell).
"""
from __future__ import annotations

import numpy as np


class CellLayout:
    """Segmentation of a cell-sorted particle array: cell ``c`` owns ``[offsets[c], offsets[c] + lengths[c])``."""

    def __init__(self, centres: np.ndarray, offsets: np.ndarray, lengths: np.ndarray):
        self._centres = np.ascontiguousarray(centres)
        self._offsets = offsets
        self._lengths = lengths
        self._num_particles = lengths.sum()
        # neighbouring centres along the fastest axis are one cell apart (cell_layout.py:15)
        self._cell_size = np.linalg.norm(self._centres[1] - self._centres[0])

    # -- orderings ------------------------------------------------------------------------------------------
    def randomize_within_cells(self):
        """A permutation of 0..N-1 that shuffles particles uniformly inside each cell and never across cells, so that
        any leading fraction of a cell is a fair subsample (cell_layout.py:17-24; unseeded, like the reference)."""
        total = int(self._lengths.sum())
        cell_of_slot = np.repeat(np.arange(len(self._lengths), dtype=_key_dtype(len(self._lengths))), self._lengths)
        # a uniform permutation of all slots, regrouped by cell with a STABLE sort: inside every cell the slots keep the
        # (uniformly random) relative order they have in the permutation.  O(N): numpy radix-sorts 16-bit keys.
        perm = np.random.permutation(total)
        return perm[np.argsort(cell_of_slot[perm], kind='stable')].astype(np.uintp)

    def cells_in_sphere(self, centre, radius: float) -> np.ndarray:
        """Indices of the cells whose centre lies within radius + one cell diagonal of ``centre`` (:26-31)."""
        reach = radius + self._cell_size * np.sqrt(3.0)
        distance = np.linalg.norm(self._centres - centre, axis=1)
        return np.where(distance < reach)[0]

    # -- lookups --------------------------------------------------------------------------------------------
    def cell_index_from_offset(self, offset: int) -> int:
        cell = np.searchsorted(self._offsets, offset, side='right') - 1
        if cell < 0 or cell >= len(self._lengths):
            raise ValueError("Offset is out of bounds")
        return cell

    def cell_slice(self, cell_index: int) -> slice:
        first = self._offsets[cell_index]
        return slice(first, first + self._lengths[cell_index])

    def get_num_cells(self):
        return len(self._lengths)

    def get_num_particles(self):
        return self._num_particles

    def get_cell_length(self, cell_index):
        return self._lengths[cell_index]

    def get_cell_offset(self, cell_index):
        return self._offsets[cell_index]

    # -- construction ---------------------------------------------------------------------------------------
    @staticmethod
    def _grid(box_min, box_max, nside):
        cell_size = (box_max - box_min) / nside
        first_centre = box_min + cell_size / 2
        axis = slice(first_centre, box_max, cell_size)
        centres = np.mgrid[axis, axis, axis].reshape(3, -1).T        # x-major, z fastest: matches iz + n(iy + n ix)
        return cell_size, centres

    @classmethod
    def from_positions(cls, particle_positions, box_min: float, box_max: float, nside: int, shuffle_seed=None):
        """Returns ``(layout, ordering)`` where ``positions[ordering]`` is cell-sorted.

        ``particle_positions`` is an (N,3) numpy array, or an (N,3) torch CUDA tensor (then the ordering is returned
        as a CUDA int64 tensor and all O(N) work runs on the device).  Raises ValueError when a particle is outside
        the box, like the reference (cell_layout.py:83-84, :100-101).

        ``shuffle_seed`` (device tensors only): fuse ``randomize_within_cells`` into the sort -- the returned ordering is
        already shuffled inside every cell (a keyed pseudo-random permutation per cell, computed slot by slot on the
        device), i.e. it equals ``ordering[layout.randomize_within_cells()]`` in distribution."""
        if _is_cuda_tensor(particle_positions):
            return cls._from_positions_device(particle_positions, box_min, box_max, nside, shuffle_seed)
        if shuffle_seed is not None:
            raise ValueError("shuffle_seed needs device-resident positions; use randomize_within_cells() on the host")

        pos = particle_positions
        if pos.min() < box_min or pos.max() >= box_max:
            raise ValueError("Particle positions are outside the box")
        cell_size, centres = cls._grid(box_min, box_max, nside)
        ijk = np.floor((pos - box_min) / cell_size).astype(np.intp)
        if ijk.min() < 0 or ijk.max() >= nside:
            raise ValueError("Particle positions are too close to edge of box; expand box size")
        cell = ijk[:, 2] + nside * (ijk[:, 1] + nside * ijk[:, 0])
        # same result as a stable argsort of the intp keys; 16-bit keys take numpy's O(N) radix sort
        ordering = np.argsort(cell.astype(_key_dtype(nside ** 3), copy=False), kind='stable')
        lengths = np.bincount(cell, minlength=nside ** 3)
        assert len(lengths) == len(centres)
        offsets = np.cumsum(lengths) - lengths
        return cls(centres, offsets, lengths), ordering

    @classmethod
    def _from_positions_device(cls, pos, box_min, box_max, nside, shuffle_seed=None):
        import ctypes

        import torch

        from . import _native as N

        if pos.dim() != 2 or pos.shape[1] != 3 or pos.dtype not in (torch.float32, torch.float64):
            raise ValueError("device positions must be an (N,3) float32/float64 tensor")
        pos = pos.contiguous()
        np_dtype = np.float32 if pos.dtype == torch.float32 else np.float64
        lo, hi = pos.min().item(), pos.max().item()
        if lo < box_min or hi >= box_max:
            raise ValueError("Particle positions are outside the box")
        # scalars exactly as numpy would form them for an array of this dtype (NEP 50: python floats are weak,
        # numpy scalars keep their own precision)
        bmin = box_min if isinstance(box_min, np.generic) else np_dtype(box_min)
        bmax = box_max if isinstance(box_max, np.generic) else np_dtype(box_max)
        if np.result_type(np_dtype, bmin.dtype) != np_dtype or np.result_type(np_dtype, bmax.dtype) != np_dtype:
            raise ValueError("box scalars of higher precision than the positions are not supported on the device path")
        cell_size, centres = cls._grid(bmin, bmax, nside)
        sub_min = np_dtype(bmin)                      # the kernel evaluates (pos - box_min) / cell_size in the position dtype
        cell_size = np_dtype(cell_size)
        n = pos.shape[0]
        lib = N.lib()
        dev = pos.device
        order = torch.empty(n, dtype=torch.int64, device=dev)
        lengths = torch.empty(nside ** 3, dtype=torch.int64, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        work = torch.empty(lib.tsplat_cell_layout_work_bytes(n, nside), dtype=torch.uint8, device=dev)
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        if shuffle_seed is None:
            N.check(lib.tsplat_cell_layout(dev.index, ctypes.c_void_p(pos.data_ptr()), n, pos.element_size(),
                                           ctypes.c_double(float(sub_min)), ctypes.c_double(float(cell_size)), nside,
                                           ctypes.c_void_p(order.data_ptr()), ctypes.c_void_p(lengths.data_ptr()),
                                           ctypes.c_void_p(status.data_ptr()), ctypes.c_void_p(work.data_ptr()),
                                           work.numel(), stream))
        else:
            N.check(lib.tsplat_cell_layout_shuffled(dev.index, ctypes.c_void_p(pos.data_ptr()), n, pos.element_size(),
                                                    ctypes.c_double(float(sub_min)), ctypes.c_double(float(cell_size)), nside,
                                                    int(shuffle_seed) & 0xffffffff or 1,
                                                    ctypes.c_void_p(order.data_ptr()), ctypes.c_void_p(lengths.data_ptr()),
                                                    ctypes.c_void_p(status.data_ptr()), ctypes.c_void_p(work.data_ptr()),
                                                    work.numel(), stream))
        if int(status.item()) != 0:
            raise ValueError("Particle positions are too close to edge of box; expand box size")
        lengths_h = lengths.cpu().numpy().astype(np.intp)
        offsets = np.cumsum(lengths_h) - lengths_h
        return cls(centres, offsets, lengths_h), order


def _key_dtype(n_cells: int):
    return np.int16 if n_cells <= 32767 else np.intp


def _is_cuda_tensor(obj) -> bool:
    try:
        import torch
    except ImportError:      # pragma: no cover
        return False
    return isinstance(obj, torch.Tensor) and obj.is_cuda
