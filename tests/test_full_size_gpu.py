"""Parity at BASELINE.json's FULL sizes (c1: 1M @ 512^2 density, c2: 10M @ 1024^2 density, c3: 50M @ 2048^2 two-channel
in EXPORT blocks, c4: 100M @ 2048^2 RGB, c5: one GPU's 125M share @ 4096^2 density) -- the synthetic workloads bench.py
times, generated the way bench.py generates them (one snapshot in topsy's cell order, synthetic.generate_striped).

Two kinds of checks:
  * the whole image against the C/OpenMP oracle (fp64 accumulators; it finishes each workload in seconds on the box's
    host cores) with the north_star tolerance: relative error <= 1e-4 wherever a pixel exceeds 1e-6 of the channel maximum;
  * size-independent properties that need no oracle image: the fragment-count channel of the RGB mode is an exact
    integer checksum of the coverage decisions (sum == number of (particle, pixel centre) pairs counted on the CPU),
    superposition (render(A) + render(B) == render(A u B)), and progressive blocks == one-shot render.
"""
import dataclasses

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import c_oracle as co
from oracle import topsy_oracle as o
from test_gpu_parity import assert_image_parity

from topsy_b200 import _native as N
from topsy_b200 import camera, synthetic
from topsy_b200.engine import SplatEngine

MODE = {"density": N.MODE_DENSITY, "weighted": N.MODE_WEIGHTED, "rgb": N.MODE_RGB}


def _setup(name):
    wl = synthetic.WORKLOADS[name]
    dev = torch.device("cuda", 0)
    data, _ = synthetic.generate_striped(wl, dev, n_total=wl.n_particles)
    wl = dataclasses.replace(wl, n_particles=int(data["x"].numel()))
    rot = camera.rotate(np.eye(3), *wl.rotate)
    M = camera.transform_matrix(rot, np.zeros(3), wl.scale); sf = np.float32(1.0 / wl.scale)
    eng = SplatEngine(wl.resolution)
    eng.set_camera(M, sf)
    eng.set_particles(data["x"], data["y"], data["z"], data["h"])
    names = synthetic.weight_names(wl.mode)
    eng.set_weights(*[data[k] for k in names])
    return wl, data, names, M, sf, eng


def _blocks(n, block=2 ** 25):
    starts = np.arange(0, n, block, dtype=np.int64)
    return starts, np.minimum(block, n - starts)


@pytest.mark.parametrize("name", ["c1", "c2", "c3", "c4", "c5"])
def test_full_size_workload_against_oracle(name, oracle_lut):
    wl, data, names, M, sf, eng = _setup(name)
    try:
        n = wl.n_particles
        mode = MODE[wl.mode]
        starts, lens = _blocks(n)
        for i, (s, l) in enumerate(zip(starts, lens)):                 # EXPORT-style blocks, like bench.py and sph.render
            img = eng.render(mode, [s], [l], clear=(i == 0))
        got = img.cpu().numpy()
        stats = eng.stats()
        assert stats["particles_submitted"] == n
        host = {k: v.cpu().numpy() for k, v in data.items()}
        ref = co.splat(host["x"], host["y"], host["z"], host["h"], [host[k] for k in names], M, sf, wl.resolution, mode, oracle_lut)
        assert_image_parity(got, ref, f"{name} full size")
        upd, culled = co.count_updates(host["x"], host["y"], host["z"], host["h"], M, sf, wl.resolution)
        assert stats["particles_culled"] >= culled                      # the kernel also counts degenerate h as culled
        if wl.mode == "rgb":
            # exact integer checksum of every coverage decision of 1e8 particles
            assert got[..., 3].max() < 2 ** 24
            assert int(got[..., 3].astype(np.float64).sum()) == upd
            assert np.array_equal(got[..., 3], ref[..., 3].astype(np.float32))
    finally:
        eng.close()


def test_superposition_and_block_independence_full_size(oracle_lut):
    """c4 at full size: odd/even interleaved ranges rendered separately add up to the one-shot render (fp32 order noise
    only), and the count channel adds up exactly."""
    wl, data, names, M, sf, eng = _setup("c4")
    try:
        n = wl.n_particles
        whole = eng.render(N.MODE_RGB).clone()
        chunk = 1 << 20
        starts = np.arange(0, n, chunk, dtype=np.int64)
        lens = np.minimum(chunk, n - starts)
        a = eng.render(N.MODE_RGB, starts[0::2], lens[0::2], clear=True).clone()
        b = eng.render(N.MODE_RGB, starts[1::2], lens[1::2], clear=True).clone()
        both = (a.double() + b.double()).cpu().numpy()
        w = whole.cpu().numpy()
        assert np.array_equal(both[..., 3], w[..., 3].astype(np.float64))
        assert_image_parity(w, both, "superposition")
        # accumulate B on top of A without clearing == whole
        eng.render(N.MODE_RGB, starts[0::2], lens[0::2], clear=True)
        acc = eng.render(N.MODE_RGB, starts[1::2], lens[1::2], clear=False).cpu().numpy()
        assert np.array_equal(acc[..., 3], w[..., 3])
        assert_image_parity(acc, both, "accumulate")
    finally:
        eng.close()
