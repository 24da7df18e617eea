"""Seeded synthetic snapshots for benchmarks and large parity runs (generator G2 "uniform").

pynbody is not available, so the arrays ``pynbody.new`` would wrap are generated directly:
  pos ~ U[-L/2, L/2)^3,  h = f * L * N_total^(-1/3) * lognormal(sigma=0.5),  m = 1/N_total,
  q (temperature) lognormal(mu = ln 1e4, sigma = 1),  rgb lognormal(sigma = 1) / N_total
f = 0.1 gives the compact "star" footprint, f = 1.0 the SPH-natural "gas"/"dm" footprint.
Generation runs on the GPU with a per-(seed, rank) torch generator so 1e8 particles take well under a second.

``generate_striped`` (round 2) produces ONE snapshot for any number of ranks, already in topsy's memory order -- sorted by
the nside^3 cells of ``CellLayout`` and uniformly shuffled inside each cell (cell_layout.py:17-24, :63-113) -- and hands a
rank exactly its per-cell stripe (``distributed.shard_indices``) without materialising the whole snapshot: every attribute of
particle i is a pure function of (SEED, attribute, i) through a counter-based integer hash, so rank g of G and a single GPU
walking all G stripes see bit-identical particles.  bench.py uses it for every GPU count; the 1-GPU case is the stripe
"all of it".
"""
from __future__ import annotations

import dataclasses
import math

import torch

BOX = 100.0          # kpc
SEED = 20240601


@dataclasses.dataclass
class Workload:
    name: str
    description: str
    n_particles: int          # per GPU
    resolution: int
    mode: str                 # 'density' | 'weighted' | 'rgb'
    h_factor: float           # f above
    scale: float = 50.0       # view half-width: the whole box
    rotate: tuple = (0.3, 0.4)

    @property
    def bytes_per_particle(self):
        return {'density': 20, 'weighted': 24, 'rgb': 28}[self.mode]

    @property
    def channels(self):
        return {'density': 1, 'weighted': 2, 'rgb': 4}[self.mode]


WORKLOADS = {
    "c1": Workload("c1", "1M gas particles, 512^2 density projection", 1_000_000, 512, "density", 1.0),
    "c2": Workload("c2", "10M dark-matter particles, 1024^2 density projection + log colormap", 10_000_000, 1024, "density", 1.0),
    "c3": Workload("c3", "50M gas particles, 2048^2 density-weighted temperature (two-channel)", 50_000_000, 2048, "weighted", 1.0),
    "c4": Workload("c4", "100M star particles, 2048^2 RGB three-band render", 100_000_000, 2048, "rgb", 0.1),
    "c4s": Workload("c4s", "100M star particles with sub-pixel footprints (h_factor 0.03), 2048^2 RGB three-band render", 100_000_000, 2048,
                    "rgb", 0.03),
    "c5": Workload("c5", "1B particles over 8 GPUs (125M per GPU), 4096^2 density projection + image sum-reduce", 125_000_000, 4096,
                   "density", 0.1),
}


def generate(workload: Workload, device, n_total: int | None = None, rank: int = 0, n: int | None = None):
    """Returns dict of float32 CUDA tensors: x, y, z, h and the weight arrays of the workload's mode."""
    n = workload.n_particles if n is None else n
    n_total = n if n_total is None else n_total
    g = torch.Generator(device=device)
    g.manual_seed(SEED + 7919 * rank)
    out = {}
    for k in "xyz":
        out[k] = (torch.rand(n, generator=g, device=device, dtype=torch.float32) - 0.5) * BOX
    h0 = workload.h_factor * BOX * n_total ** (-1.0 / 3.0)
    out["h"] = torch.exp(torch.randn(n, generator=g, device=device, dtype=torch.float32) * 0.5) * h0
    if workload.mode == "rgb":
        for k in ("r", "g", "b"):
            out[k] = torch.exp(torch.randn(n, generator=g, device=device, dtype=torch.float32)) / n_total
    else:
        out["m"] = torch.full((n,), 1.0 / n_total, device=device, dtype=torch.float32)
        if workload.mode == "weighted":
            out["q"] = torch.exp(torch.randn(n, generator=g, device=device, dtype=torch.float32) + math.log(1e4))
    return out


def weight_names(mode: str):
    return {"density": ("m",), "weighted": ("m", "q"), "rgb": ("r", "g", "b")}[mode]


# ----------------------------------------------------------------------------------------------------------------
# one snapshot in cell order, striped over ranks
# ----------------------------------------------------------------------------------------------------------------
NSIDE = 16           # config.DEFAULT_CELLS_NSIDE of the reference (config.py:27)


def cell_lengths(n_total: int, nside: int = NSIDE) -> torch.Tensor:
    """Particles per cell of the synthetic uniform snapshot: equal shares, the remainder goes to the first cells."""
    ncells = nside ** 3
    base, rem = divmod(int(n_total), ncells)
    out = torch.full((ncells,), base, dtype=torch.int64)
    out[:rem] += 1
    return out


def _hash_uniform(idx: torch.Tensor, stream: int) -> torch.Tensor:
    """Counter-based U[0,1): two rounds of a 32-bit integer hash of (particle index, stream), 24 random bits."""
    m = 0xFFFFFFFF
    lo = idx & m
    hi = idx >> 32
    x = (lo * 0x9E3779B1 + hi * 0x85EBCA77 + (SEED + 0x632BE5AB * (stream + 1))) & m
    for _ in range(2):
        x = x ^ (x >> 16)
        x = (x * 0x7FEB352D) & m
        x = x ^ (x >> 15)
        x = (x * 0x846CA68B) & m
        x = x ^ (x >> 16)
    return (x >> 8).to(torch.float32) * (1.0 / 16777216.0)


def _hash_normal(idx: torch.Tensor, stream: int) -> torch.Tensor:
    u1 = _hash_uniform(idx, stream).clamp_min(2.0 ** -24)
    u2 = _hash_uniform(idx, stream + 1)
    return torch.sqrt(-2.0 * torch.log(u1)) * torch.cos(2.0 * math.pi * u2)


def stripe_size(n_total: int, rank: int, world: int, nside: int = NSIDE) -> int:
    lengths = cell_lengths(n_total, nside)
    return int(torch.clamp((lengths - rank + world - 1) // world, min=0).sum())


def generate_striped(workload: Workload, device, n_total: int, rank: int = 0, world: int = 1, nside: int = NSIDE,
                     chunk: int = 1 << 24, h_count: int | None = None):
    """Rank ``rank``'s per-cell stripe of the ``n_total``-particle snapshot (see the module docstring).

    Returns (data, lengths): data = dict of float32 tensors (x, y, z, h + the weight arrays of the workload's mode) in
    cell order, lengths = the rank's particles per cell (int64, CPU) -- what RenderProgressionWithCells would work with.
    ``h_count``: particle count the smoothing-length scale h0 = f L h_count^(-1/3) is set from (default n_total); weak-scaling
    runs pass the per-GPU count so that the footprint in pixels does not shrink as GPUs are added.
    """
    device = torch.device(device)
    lengths = cell_lengths(n_total, nside)
    offsets = torch.cumsum(lengths, 0) - lengths
    mine = torch.clamp((lengths - rank + world - 1) // world, min=0)          # distributed.shard_cell_lengths
    n = int(mine.sum())
    names = ["x", "y", "z", "h"] + list(weight_names(workload.mode))
    out = {k: torch.empty(n, dtype=torch.float32, device=device) for k in names}
    first = (torch.cumsum(mine, 0) - mine).to(device)
    mine_d, offsets_d = mine.to(device), offsets.to(device)
    cell_size = BOX / nside
    h0 = workload.h_factor * BOX * (n_total if h_count is None else h_count) ** (-1.0 / 3.0)
    cells = torch.arange(nside ** 3, device=device)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        local = torch.arange(s, e, device=device, dtype=torch.int64)
        cell = torch.searchsorted(first, local, right=True) - 1
        # empty cells share their `first` with the next cell: searchsorted(right) - 1 lands on the LAST of them, which is
        # the non-empty one only if empties are skipped -- walk back is unnecessary because equal shares leave no empty
        # cell unless n_total < ncells * world; guard it anyway
        if bool((mine_d[cell] == 0).any()):
            nonempty = cells[mine_d > 0]
            cell = nonempty[torch.searchsorted(first[mine_d > 0], local, right=True) - 1]
        gidx = offsets_d[cell] + rank + (local - first[cell]) * world             # distributed.shard_indices
        iz = cell % nside
        iy = (cell // nside) % nside
        ix = cell // (nside * nside)                                             # cell = iz + n (iy + n ix)  (cell_layout.py:95)
        for k, (axis, icell) in enumerate((("x", ix), ("y", iy), ("z", iz))):
            # u <= 1 - 2^-18 keeps icell + u (fp32 spacing 2^-20 below 16) and its image in the box strictly inside cell icell
            u = _hash_uniform(gidx, k).clamp_max(1.0 - 2.0 ** -18)
            out[axis][s:e] = (icell.to(torch.float32) + u) * cell_size - 0.5 * BOX
        out["h"][s:e] = torch.exp(_hash_normal(gidx, 3) * 0.5) * h0
        if workload.mode == "rgb":
            for k, name in enumerate(("r", "g", "b")):
                out[name][s:e] = torch.exp(_hash_normal(gidx, 5 + 2 * k)) / n_total
        else:
            out["m"][s:e] = 1.0 / n_total
            if workload.mode == "weighted":
                out["q"][s:e] = torch.exp(_hash_normal(gidx, 5) + math.log(1e4))
    return out, mine
