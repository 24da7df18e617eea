"""Multi-GPU SPH projection: particles shard, images sum (no reference counterpart -- topsy is single-device).

One process per GPU (torchrun).  Splatting is a sum over particles, so each rank splats its shard into a full-resolution
partial image and the only exchange is one image sum before the colormap:

  * sharding    per-cell striping -- rank g takes every G-th particle of every cell (``shard_indices``).  Because the
                within-cell order is a uniform shuffle, every rank holds a statistically identical subsample, so load is
                balanced for any zoom, and cell selection / progressive fractions apply unchanged with the rank's own
                per-cell lengths (``shard_cell_lengths``).
  * reduce      'p2p'   (default on NVLink boxes): the partial images live in PyTorch symmetric memory; after a
                        device-side barrier every rank runs ONE kernel (tsplat_reduce_colormap) that loads its slab of
                        rows from every peer over NVLink, adds them in rank order, applies the colormap and stores the
                        RGBA rows straight into rank 0's output -- reduce-scatter + colormap + gather fused.
                'nccl'  baseline / fallback: ``dist.all_reduce`` (or ``reduce`` to rank 0) of the fp32 image, then the
                        ordinary colormap kernel on rank 0.
                'gloo'  CPU tensors, used by the world_size-2 CPU tests of the host logic only.
"""
from __future__ import annotations

import ctypes

import numpy as np


# ----------------------------------------------------------------------------------------------------------------
# sharding arithmetic (pure numpy: tested on CPU)
# ----------------------------------------------------------------------------------------------------------------
def shard_cell_lengths(lengths: np.ndarray, rank: int, world: int) -> np.ndarray:
    """Particles of each cell that land on ``rank`` under per-cell striping: ceil((len_c - rank) / world)."""
    lengths = np.asarray(lengths, dtype=np.int64)
    return np.maximum(0, (lengths - rank + world - 1) // world)


def shard_indices(offsets: np.ndarray, lengths: np.ndarray, rank: int, world: int) -> np.ndarray:
    """Global indices (into the cell-sorted, shuffled order) owned by ``rank``: offset_c + rank + k*world."""
    offsets = np.asarray(offsets, dtype=np.int64)
    mine = shard_cell_lengths(lengths, rank, world)
    total = int(mine.sum())
    if total == 0:
        return np.zeros(0, dtype=np.int64)
    cell_of = np.repeat(np.arange(len(mine)), mine)
    first = np.cumsum(mine) - mine
    k = np.arange(total, dtype=np.int64) - first[cell_of]
    return offsets[cell_of] + rank + k * world


def row_slab(resolution: int, rank: int, world: int):
    """Rows [row0, row0 + nrows) of the image that ``rank`` reduces and colormaps."""
    base, extra = divmod(resolution, world)
    row0 = rank * base + min(rank, extra)
    return row0, base + (1 if rank < extra else 0)


# NVLink cost of a byte of RGBA that a rank stores into (or receives from) a peer, relative to a byte of partial image it
# LOADS from a peer: loads are round trips, stores are posted.  Measured with the fused kernel on 2 x B200: a rank that
# only loads 67 MB takes 0.149 ms (452 GB/s); with equal slabs (33.5 MB loaded per rank, 33.5 MB of RGBA crossing once)
# both ranks take 0.113 ms, i.e. a stored byte costs its sender and its receiver ~0.53 of a loaded byte.
REMOTE_STORE_COST = 0.53


def presentation_slab(resolution: int, rank: int, world: int, in_bytes_per_pixel: int, out_bytes_per_pixel: int,
                      dst: int = 0):
    """Rows of the image that ``rank`` reduces and colormaps when the RGBA result is gathered on rank ``dst``.

    What bounds the fused kernel is each GPU's NVLink work.  A rank that owns n rows loads them from the other world - 1
    partial images and (unless it is ``dst``) stores n rows of RGBA to ``dst``; ``dst`` receives the RGBA rows of
    everybody else.  With equal slabs ``dst`` is the straggler on more than two GPUs (c5 on 8 GPUs: 59 MB loaded and 59 MB
    received, 0.195 ms, against 59 MB + 8 MB sent for the others); its slab is chosen so that the slowest rank is as fast
    as possible (two GPUs: equal slabs; c5 on 8: 7 % of the rows instead of 12.5 %)."""
    if world == 1:
        return 0, resolution
    i_b, o_b = float(in_bytes_per_pixel), REMOTE_STORE_COST * float(out_bytes_per_pixel)
    # exact over the integer row counts (covers the corner where the output is larger than a partial image, too)
    n = np.arange(resolution + 1, dtype=np.float64)
    cost_dst = (world - 1) * n * i_b + (resolution - n) * o_b
    cost_others = np.ceil((resolution - n) / (world - 1)) * ((world - 1) * i_b + o_b)
    n_dst = int(np.argmin(np.maximum(cost_dst, cost_others)))
    others = [r for r in range(world) if r != dst]
    base, extra = divmod(resolution - n_dst, world - 1)
    rows = {r: base + (1 if k < extra else 0) for k, r in enumerate(others)}
    rows[dst] = n_dst
    row0 = sum(rows[r] for r in range(rank))
    return row0, rows[rank]


# ----------------------------------------------------------------------------------------------------------------
# sharding of the drop-in classes (Visualizer / loaders / SPH) under torchrun
# ----------------------------------------------------------------------------------------------------------------
_sharding_enabled = True


def set_sharding(enabled: bool):
    """Switch the automatic sharding of the drop-in classes on / off (e.g. to build an unsharded reference Visualizer
    inside a multi-rank test).  Returns the previous setting."""
    global _sharding_enabled
    previous, _sharding_enabled = _sharding_enabled, bool(enabled)
    return previous


def shard_context(group=None):
    """(rank, world) the drop-in classes shard over: the default process group when torch.distributed is initialised
    with more than one rank and sharding is enabled, else (0, 1)."""
    if not _sharding_enabled:
        return 0, 1
    try:
        import torch.distributed as dist
    except ImportError:      # pragma: no cover
        return 0, 1
    if not (dist.is_available() and dist.is_initialized()):
        return 0, 1
    world = dist.get_world_size(group)
    return (dist.get_rank(group), world) if world > 1 else (0, 1)


def shard_loader(loader, rank: int, world: int):
    """Keep only ``rank``'s stripe of a data loader: every ``world``-th particle of every cell of its cell layout (or of
    the whole snapshot when the loader has no cells).  The loader keeps working as before -- ``len``, the getters and
    ``get_render_progression`` now describe the stripe -- and remembers the global particle count in
    ``loader.global_num_particles``."""
    from .cell_layout import CellLayout
    n = len(loader)
    if world <= 1:
        loader.global_num_particles = n
        return loader
    layout = getattr(loader, "_cell_layout", None)
    if layout is not None:
        mine = shard_indices(layout._offsets, layout._lengths, rank, world)
        lengths = shard_cell_lengths(layout._lengths, rank, world).astype(layout._lengths.dtype)
        loader._cell_layout = CellLayout(layout._centres, np.cumsum(lengths) - lengths, lengths)
    else:
        mine = np.arange(rank, n, world, dtype=np.int64)
    loader._keep_particles(mine)
    loader.global_num_particles = n
    return loader


def reduce_image_host(image, group=None, dst: int | None = None):
    """Sum a partial image over the process group with the backend's collective (NCCL on CUDA tensors, gloo on CPU)."""
    import torch.distributed as dist
    if dst is None:
        dist.all_reduce(image, op=dist.ReduceOp.SUM, group=group)
    else:
        dist.reduce(image, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return image


# ----------------------------------------------------------------------------------------------------------------
# device side
# ----------------------------------------------------------------------------------------------------------------
class ImageExchange:
    """Partial + reduced accumulation image of one renderer on every rank, and the all-reduce between them.

    'p2p'  (NVLink boxes): both images and a one-float mass-scale slot live in PyTorch symmetric memory; ``allreduce`` is
           barrier -> tsplat_allreduce_image (every rank reduces its slab of rows over all peers' partial images, weighted
           by each peer's mass scale, and stores it into every peer's reduced image) -> barrier.
    'nccl' fallback where peer access is not available: all_reduce of scale * partial.
    Construction is collective: every rank must create its exchanges in the same order."""

    def __init__(self, engine, resolution: int, channels: int, group=None, method: str = "auto"):
        import torch
        import torch.distributed as dist
        self.engine = engine
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.resolution, self.channels = resolution, channels
        self.device = engine.device
        shape = (resolution, resolution, channels)
        self.method = method
        if method in ("auto", "p2p"):
            try:
                import torch.distributed._symmetric_memory as symm
                gname = (group or dist.group.WORLD).group_name
                self.partial = symm.empty(shape, dtype=torch.float32, device=self.device)
                self.reduced = symm.empty(shape, dtype=torch.float32, device=self.device)
                self._scale = symm.empty((4,), dtype=torch.float32, device=self.device)
                self._hdl = symm.rendezvous(self.partial, gname)
                hdl_r = symm.rendezvous(self.reduced, gname)
                hdl_s = symm.rendezvous(self._scale, gname)
                self.partial.zero_(); self.reduced.zero_(); self._scale.fill_(1.0)
                arr = ctypes.c_void_p * self.world
                self._peer_partial = arr(*[int(p) for p in self._hdl.buffer_ptrs])
                self._peer_reduced = arr(*[int(p) for p in hdl_r.buffer_ptrs])
                self._peer_scale = arr(*[int(p) for p in hdl_s.buffer_ptrs])
                self.method = "p2p"
            except Exception as e:
                if method == "p2p":
                    raise
                self.method = "nccl"
                self._fallback_reason = repr(e)
        if self.method == "nccl":
            self.partial = torch.zeros(shape, dtype=torch.float32, device=self.device)
            self.reduced = torch.zeros(shape, dtype=torch.float32, device=self.device)

    def allreduce(self, mass_scale: float, zmax: bool = False):
        """reduced (on every rank) = sum over ranks of mass_scale_r * partial_r   (zmax: per-pixel z-buffer maximum)."""
        import torch
        import torch.distributed as dist
        from . import _native as N
        if self.method == "p2p":
            self._scale.fill_(float(mass_scale))
            self._hdl.barrier()                       # every partial image and scale is complete and visible
            row0, nrows = row_slab(self.resolution, self.rank, self.world)
            stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            N.check(self.engine.lib.tsplat_allreduce_image(self.engine._ctx, self._peer_partial, self._peer_reduced,
                                                           self._peer_scale, self.world, self.channels, row0, nrows,
                                                           N.REDUCE_ZMAX if zmax else N.REDUCE_SUM, stream))
            self._hdl.barrier()                       # every slab of every reduced image has been stored
        elif zmax:
            keys = self.partial.view(torch.int64).clone()     # (quantity, depth) pixels as 64-bit keys, depth high
            dist.all_reduce(keys, op=dist.ReduceOp.MAX, group=self.group)
            self.reduced.copy_(keys.view(torch.float32).view_as(self.reduced))
        else:
            torch.mul(self.partial, float(mass_scale), out=self.reduced)
            dist.all_reduce(self.reduced, op=dist.ReduceOp.SUM, group=self.group)
        return self.reduced


class ShardedSplat:
    """One rank of a sharded render: local engine + the image exchange."""

    def __init__(self, resolution: int, channels: int, out_format: str = "rgba8unorm", reduce: str = "auto", group=None):
        import torch
        import torch.distributed as dist

        from . import _native as N
        from .engine import SplatEngine

        self.N = N
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.group = group
        self.resolution = resolution
        self.channels = channels
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.engine = SplatEngine(resolution, device=self.device.index)
        self.out_format = out_format
        fmt, tdtype = {"rgba8unorm": (N.FMT_RGBA8, torch.uint8), "rgba16float": (N.FMT_RGBA16F, torch.float16),
                       "rgba32float": (N.FMT_RGBA32F, torch.float32)}[out_format]
        self._fmt = fmt
        self._out_pixel_bytes = {N.FMT_RGBA8: 4, N.FMT_RGBA16F: 8, N.FMT_RGBA32F: 16}[fmt]
        self.method = reduce
        if reduce == "auto":
            self.method = "p2p" if self.world > 1 else "local"
        self._hdl = self._out_hdl = None
        self.kernel_events = None          # optional (start, end) CUDA events around the fused kernel alone (bench.py)
        shape = (resolution, resolution, channels)
        if self.method == "p2p" and self.world > 1:
            try:
                import torch.distributed._symmetric_memory as symm
                self.image = symm.empty(shape, dtype=torch.float32, device=self.device)
                self.out = symm.empty((resolution, resolution, 4), dtype=tdtype, device=self.device)
                gname = (group or dist.group.WORLD).group_name
                self._hdl = symm.rendezvous(self.image, gname)
                self._out_hdl = symm.rendezvous(self.out, gname)
                self.image.zero_()
                self._peer_ptrs = (ctypes.c_void_p * self.world)(*[int(p) for p in self._hdl.buffer_ptrs])
                self._out0 = int(self._out_hdl.buffer_ptrs[0])
            except Exception as e:      # no peer access (e.g. PCIe-only box): fall back to the collective
                if reduce == "p2p":
                    raise
                self.method = "nccl"
                self._fallback_reason = repr(e)
        if self.method != "p2p" or self.world == 1:
            self.image = torch.zeros(shape, dtype=torch.float32, device=self.device)
            self.out = torch.empty((resolution, resolution, 4), dtype=tdtype, device=self.device)
        self.engine.bind_image(self.image)

    # -- frame ------------------------------------------------------------------------------------------------
    def splat(self, mode, blocks):
        """Local splat of this rank's particles (already set on ``self.engine``): list of (start, length) blocks."""
        for i, (s, l) in enumerate(blocks):
            self.engine.render(mode, [s], [l], clear=(i == 0), image=self.image)

    def present(self, params, lut):
        """Image sum over ranks + colormap.  The RGBA result is valid on rank 0 (``self.out``)."""
        import torch.distributed as dist
        N = self.N
        eng = self.engine
        if self.world == 1 or self.method == "local":
            eng.colormap(self.image, params, lut, self.out, self._fmt)
            return self.out
        if self.method == "p2p":
            import torch
            self._hdl.barrier()                                   # every rank's partial image is complete and visible
            if self.kernel_events is not None:
                self.kernel_events[0].record()
            row0, nrows = presentation_slab(self.resolution, self.rank, self.world, 4 * self.channels, self._out_pixel_bytes)
            lw, lh = (0, 0) if lut is None else ((lut.shape[0], 1) if lut.dim() == 2 else (lut.shape[1], lut.shape[0]))
            stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            N.check(eng.lib.tsplat_reduce_colormap(eng._ctx, self._peer_ptrs, self.world, self.channels, row0, nrows,
                                                   ctypes.byref(params), None if lut is None else ctypes.c_void_p(lut.data_ptr()),
                                                   lw, lh, ctypes.c_void_p(self._out0), self._fmt, None, stream))
            if self.kernel_events is not None:
                self.kernel_events[1].record()
            self._out_hdl.barrier()                               # rank 0's output is complete; images may be cleared
            return self.out
        # nccl baseline: reduce a copy so that the partial image stays valid (progressive REFINE frames keep adding to it)
        if getattr(self, "_sum", None) is None:
            self._sum = self.image.clone()
        else:
            self._sum.copy_(self.image)
        dist.reduce(self._sum, dst=0, op=dist.ReduceOp.SUM, group=self.group)
        if self.rank == 0:
            eng.colormap(self._sum, params, lut, self.out, self._fmt)
        return self.out

    def reduced_image(self):
        """fp32 sum image on every rank (for get_image-style readback and tests); collective."""
        import torch.distributed as dist
        img = self.image.clone()
        if self.world > 1:
            dist.all_reduce(img, op=dist.ReduceOp.SUM, group=self.group)
        return img

    def close(self):
        self.engine.close()
