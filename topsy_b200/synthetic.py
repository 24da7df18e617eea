"""Seeded synthetic snapshots for benchmarks and large parity runs (generator G2 "uniform").

pynbody is not available, so the arrays ``pynbody.new`` would wrap are generated directly:
  pos ~ U[-L/2, L/2)^3,  h = f * L * N_total^(-1/3) * lognormal(sigma=0.5),  m = 1/N_total,
  q (temperature) lognormal(mu = ln 1e4, sigma = 1),  rgb lognormal(sigma = 1) / N_total
f = 0.1 gives the compact "star" footprint, f = 1.0 the SPH-natural "gas"/"dm" footprint.
Generation runs on the GPU with a per-(seed, rank) torch generator so 1e8 particles take well under a second.
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
