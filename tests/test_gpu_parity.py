"""GPU parity: the CUDA splat (through the C ABI) against the CPU oracle on identical seeded inputs.

Tolerance (BASELINE.json north_star): per-pixel relative error <= 1e-4 on every accumulated channel wherever the
oracle pixel exceeds 1e-6 of the channel's maximum.  The oracle accumulates in fp64; the GPU in fp32 atomics.
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import c_oracle as co
from oracle import topsy_oracle as o

REL_TOL = 1e-4
FLOOR = 1e-6


def assert_image_parity(gpu, ref, what="", mag=None):
    """|gpu - ref| <= 1e-4 * mag wherever mag > 1e-6 * max(mag).  ``mag`` is the oracle image of the same particles
    with |weights| (== ref for the positive-definite channels: density, RGB, count, depth), i.e. the plain relative
    error of the north_star for everything except a *signed* quantity channel, where a sum that cancels has no
    meaningful relative error and the bound is taken relative to the accumulated magnitude instead."""
    gpu = np.asarray(gpu, np.float64)
    assert gpu.shape == ref.shape, (gpu.shape, ref.shape)
    mag = np.abs(ref) if mag is None else np.abs(mag)
    for c in range(ref.shape[2]):
        r = ref[..., c]; g = gpu[..., c]; a = mag[..., c]
        big = a > FLOOR * a.max()
        if big.any():
            rel = np.abs(g[big] - r[big]) / a[big]
            assert rel.max() <= REL_TOL, f"{what} channel {c}: max rel err {rel.max():.3e} at {np.argmax(rel)}"
        small = ~big
        if small.any():
            assert np.abs(g[small] - r[small]).max() <= 2 * FLOOR * max(a.max(), 1e-300), f"{what} ch {c} floor"


def oracle_pair(x, y, z, h, w, M, sf, R, mode, lut, ranges=None):
    """(reference image, magnitude image) -- the latter only differs for a signed second weight."""
    ref = co.splat(x, y, z, h, w, M, sf, R, mode, lut, ranges=ranges)
    if mode == o.MODE_WEIGHTED and (np.asarray(w[1]) < 0).any():
        mag = co.splat(x, y, z, h, (w[0], np.abs(w[1])), M, sf, R, mode, lut, ranges=ranges)
    else:
        mag = ref
    return ref, mag


@pytest.fixture(scope="module")
def engine200():
    from topsy_b200.engine import SplatEngine
    eng = SplatEngine(200)
    yield eng
    eng.close()


def _to_dev(*arrs):
    return [torch.from_numpy(np.ascontiguousarray(a, np.float32)).cuda() for a in arrs]


def _weights_for(mode, fx):
    if mode == o.MODE_RGB:
        rgb = fx.rgb
        return (rgb[:, 0], rgb[:, 1], rgb[:, 2])
    if mode == o.MODE_WEIGHTED:
        return (fx.mass, fx.quantity.astype(np.float32))
    return (fx.mass,)


@pytest.mark.parametrize("mode", [o.MODE_DENSITY, o.MODE_WEIGHTED, o.MODE_RGB, o.MODE_DEPTH])
@pytest.mark.parametrize("scale,angles", [(200.0, (0.0, 0.0)), (20.0, (0.0, 0.4)), (2.0, (0.3, 0.4)), (0.3, (1.0, -0.7)),
                                          (2000.0, (0.2, 0.1))])
def test_gmm_fixture_parity(engine200, oracle_lut, mode, scale, angles):
    fx = o.GMMFixture(4000)
    ps = fx.pos_smooth()
    rot = o.rotate(np.eye(3), *angles)
    M = o.transform_matrix(rot, np.array([0.3, -0.2, 0.1]), scale); sf = o.scale_factor(scale)
    w = _weights_for(mode, fx)
    ref, mag = oracle_pair(ps[:, 0], ps[:, 1], ps[:, 2], ps[:, 3], w, M, sf, 200, mode, oracle_lut)
    x, y, z, h = _to_dev(ps[:, 0], ps[:, 1], ps[:, 2], ps[:, 3])
    wd = _to_dev(*w)
    engine200.set_kernel_lut(oracle_lut)
    engine200.set_camera(M, sf)
    engine200.set_particles(x, y, z, h)
    engine200.set_weights(*wd)
    img = engine200.render(mode).cpu().numpy()
    assert_image_parity(img, ref, f"mode {mode} scale {scale}", mag)
    st = engine200.stats()
    assert st["particles_submitted"] == 4000
    upd, culled = co.count_updates(ps[:, 0], ps[:, 1], ps[:, 2], ps[:, 3], M, sf, 200)
    assert st["particles_culled"] == culled


def test_ranges_and_accumulate(engine200, oracle_lut):
    """Ragged, unaligned, empty ranges; clear=False accumulates like LoadOp.load (sph.py:337-349)."""
    rs = np.random.RandomState(5)
    n = 10007
    pos = rs.uniform(-1, 1, (n, 3)).astype(np.float32)
    h = (0.02 * np.exp(rs.normal(size=n) * 0.8)).astype(np.float32)
    m = rs.uniform(0.5, 1.5, n).astype(np.float32); q = rs.normal(size=n).astype(np.float32)
    M = o.transform_matrix(o.rotate(np.eye(3), 0.5, 0.2), np.zeros(3), 1.0); sf = o.scale_factor(1.0)
    starts = np.array([0, 5, 5, 1001, 4099, 9000, 10006], np.int64)
    lens = np.array([3, 0, 990, 2001, 1, 1000, 1], np.int64)
    ref, mag = oracle_pair(pos[:, 0], pos[:, 1], pos[:, 2], h, (m, q), M, sf, 200, o.MODE_WEIGHTED, oracle_lut, ranges=(starts, lens))
    ref2, mag2 = oracle_pair(pos[:, 0], pos[:, 1], pos[:, 2], h, (m, q), M, sf, 200, o.MODE_WEIGHTED, oracle_lut,
                             ranges=(np.array([6000], np.int64), np.array([2500], np.int64)))
    x, y, z, hd, md, qd = _to_dev(pos[:, 0], pos[:, 1], pos[:, 2], h, m, q)
    engine200.set_kernel_lut(oracle_lut)
    engine200.set_camera(M, sf)
    engine200.set_particles(x, y, z, hd)
    engine200.set_weights(md, qd)
    img = engine200.render(o.MODE_WEIGHTED, starts, lens).cpu().numpy()
    assert_image_parity(img, ref, "ranges", mag)
    assert engine200.stats()["particles_submitted"] == lens.sum()
    img = engine200.render(o.MODE_WEIGHTED, [6000], [2500], clear=False).cpu().numpy()
    assert_image_parity(img, ref + ref2, "accumulate", mag + mag2)
    with pytest.raises(ValueError):
        engine200.render(o.MODE_WEIGHTED, [10000], [8])


def test_empty_and_degenerate(engine200, oracle_lut):
    """Zero particles, zero / negative / NaN smoothing lengths, particles far off screen."""
    x, y, z, h, m = _to_dev(np.zeros(0), np.zeros(0), np.zeros(0), np.zeros(0), np.zeros(0))
    engine200.set_camera(o.transform_matrix(np.eye(3), np.zeros(3), 1.0), 1.0)
    engine200.set_particles(x, y, z, h); engine200.set_weights(m)
    assert float(engine200.render(o.MODE_DENSITY).abs().sum()) == 0.0
    pos = np.array([[0, 0, 0], [0.1, 0.1, 0], [0.2, 0, 0], [1e6, 0, 0], [0, -1e6, 0], [0.5, 0.5, 0.0], [0, 0, 5.0]], np.float32)
    hh = np.array([0.0, -1.0, np.nan, 0.1, 0.1, 0.05, 0.1], np.float32)
    mm = np.ones(len(hh), np.float32)
    M = o.transform_matrix(np.eye(3), np.zeros(3), 1.0)
    ref = co.splat(pos[:, 0], pos[:, 1], pos[:, 2], hh, (mm,), M, 1.0, 200, o.MODE_DENSITY, oracle_lut)
    x, y, z, h, m = _to_dev(pos[:, 0], pos[:, 1], pos[:, 2], hh, mm)
    engine200.set_particles(x, y, z, h); engine200.set_weights(m)
    img = engine200.render(o.MODE_DENSITY).cpu().numpy()
    assert np.isfinite(img).all()
    assert_image_parity(img, ref, "degenerate")


# the last four cases defer > 131072 records, i.e. they go through the tile binning and the column-strip gather (K2 + K3)
# in every mode, with a mix of nearest-texel (< 64 px) and bilinear (>= 64 px) footprints
@pytest.mark.parametrize("mode,R,n,hscale", [(o.MODE_DENSITY, 512, 1_000_000, 0.002), (o.MODE_RGB, 1024, 2_000_000, 0.001),
                                            (o.MODE_WEIGHTED, 256, 200_000, 0.05), (o.MODE_RGB, 256, 200_000, 0.1),
                                            (o.MODE_DEPTH, 320, 200_000, 0.08), (o.MODE_DENSITY, 200, 200_000, 0.3)])
def test_uniform_box_parity(oracle_lut, mode, R, n, hscale):
    """Larger seeded runs (the oracle's C/OpenMP restatement finishes in seconds)."""
    from topsy_b200.engine import SplatEngine
    rs = np.random.RandomState(11)
    pos = rs.uniform(-1, 1, (n, 3)).astype(np.float32)
    h = (hscale * np.exp(rs.normal(size=n) * 0.5)).astype(np.float32)
    w = [rs.uniform(0.1, 1.0, n).astype(np.float32) for _ in range(3)]
    w = {o.MODE_DENSITY: w[:1], o.MODE_WEIGHTED: w[:2], o.MODE_RGB: w[:3], o.MODE_DEPTH: w[:1]}[mode]
    M = o.transform_matrix(o.rotate(np.eye(3), 0.3, 0.4), np.zeros(3), 1.0); sf = o.scale_factor(1.0)
    ref = co.splat(pos[:, 0], pos[:, 1], pos[:, 2], h, w, M, sf, R, mode, oracle_lut)
    eng = SplatEngine(R)
    try:
        eng.set_kernel_lut(oracle_lut)
        eng.set_camera(M, sf)
        x, y, z, hd = _to_dev(pos[:, 0], pos[:, 1], pos[:, 2], h)
        eng.set_particles(x, y, z, hd); eng.set_weights(*_to_dev(*w))
        img = eng.render(mode).cpu().numpy()
        assert_image_parity(img, ref, f"uniform mode {mode}")
    finally:
        eng.close()


# Seeded sweep over what the fixed cases above keep constant: resolutions that are not multiples of the 128-bit cell width
# (density / weighted fall back to one-pixel cells), views that cut the particle cloud, offsets, signed second weights,
# ragged ranges, every footprint regime from sub-pixel to bilinear in one image, and every mode.
@pytest.mark.parametrize("seed,mode,R,n,px_med", [
    (0, o.MODE_DENSITY, 50, 3000, 0.6), (1, o.MODE_WEIGHTED, 97, 20000, 2.5), (2, o.MODE_RGB, 201, 50000, 1.2),
    (3, o.MODE_DEPTH, 333, 40000, 4.0), (4, o.MODE_DENSITY, 201, 250000, 12.0), (5, o.MODE_WEIGHTED, 333, 220000, 9.0),
    (6, o.MODE_RGB, 97, 180000, 14.0), (7, o.MODE_DEPTH, 201, 200000, 30.0), (8, o.MODE_DENSITY, 1000, 300000, 0.3),
    (9, o.MODE_WEIGHTED, 640, 1, 40.0), (10, o.MODE_RGB, 512, 37, 90.0), (11, o.MODE_DENSITY, 333, 150000, 70.0)])
def test_seeded_configuration_sweep(oracle_lut, seed, mode, R, n, px_med):
    from topsy_b200.engine import SplatEngine
    rs = np.random.RandomState(4200 + seed)
    scale = float(10 ** rs.uniform(-0.3, 1.3))
    pos = (rs.normal(size=(n, 3)) * scale * rs.uniform(0.4, 1.6)).astype(np.float32)      # part of the cloud is out of view
    h = (px_med * scale / (2.0 * R) * np.exp(rs.normal(size=n) * 0.8)).astype(np.float32)  # wpx = 2 h R / scale
    m = rs.uniform(0.1, 1.0, n).astype(np.float32)
    w = {o.MODE_DENSITY: (m,), o.MODE_DEPTH: (m,), o.MODE_WEIGHTED: (m, rs.normal(size=n).astype(np.float32)),
         o.MODE_RGB: (m, rs.uniform(0, 1, n).astype(np.float32), rs.uniform(0, 2, n).astype(np.float32))}[mode]
    rot = o.rotate(np.eye(3), rs.uniform(-3, 3), rs.uniform(-1.5, 1.5))
    M = o.transform_matrix(rot, rs.normal(size=3) * 0.2 * scale, scale); sf = o.scale_factor(scale)
    ranges = None
    if seed % 3 == 1 and n > 100:                       # ragged, unaligned ranges with gaps
        cuts = np.sort(rs.choice(n, size=8, replace=False))
        starts = cuts[::2].astype(np.int64); lens = (cuts[1::2] - cuts[::2]).astype(np.int64)
        ranges = (starts, lens)
    ref, mag = oracle_pair(pos[:, 0], pos[:, 1], pos[:, 2], h, w, M, sf, R, mode, oracle_lut, ranges=ranges)
    eng = SplatEngine(R)
    try:
        eng.set_kernel_lut(oracle_lut)
        eng.set_camera(M, sf)
        x, y, z, hd = _to_dev(pos[:, 0], pos[:, 1], pos[:, 2], h)
        eng.set_particles(x, y, z, hd); eng.set_weights(*_to_dev(*w))
        img = (eng.render(mode) if ranges is None else eng.render(mode, ranges[0], ranges[1])).cpu().numpy()
        assert_image_parity(img, ref, f"sweep seed {seed} mode {mode} R {R}", mag)
        want = n if ranges is None else int(ranges[1].sum())
        assert eng.stats()["particles_submitted"] == want
    finally:
        eng.close()


def test_image_wider_than_the_direct_path_packs(oracle_lut):
    """R > 8192: K1 packs pixel coordinates into 13 bits, so on larger images every covered particle takes the deferred
    route (tsplat_project.cuh, KP_MAX_R).  Few particles, one oracle thread: the fp64 reference image is 0.5 GB."""
    from topsy_b200.engine import SplatEngine
    R, n = 8200, 3000
    rs = np.random.RandomState(77)
    pos = rs.uniform(-1, 1, (n, 3)).astype(np.float32)
    h = (np.array([0.3, 2.0, 9.0, 40.0])[rs.randint(4, size=n)] / (2.0 * R) * np.exp(rs.normal(size=n) * 0.3)).astype(np.float32)
    m = rs.uniform(0.1, 1.0, n).astype(np.float32)
    M = o.transform_matrix(np.eye(3), np.zeros(3), 1.0); sf = o.scale_factor(1.0)
    ref = co.splat(pos[:, 0], pos[:, 1], pos[:, 2], h, (m,), M, sf, R, o.MODE_DENSITY, oracle_lut, nthreads=1)
    eng = SplatEngine(R, max_particles_per_call=1 << 16)
    try:
        eng.set_kernel_lut(oracle_lut)
        eng.set_camera(M, sf)
        x, y, z, hd, md = _to_dev(pos[:, 0], pos[:, 1], pos[:, 2], h, m)
        eng.set_particles(x, y, z, hd); eng.set_weights(md)
        img = eng.render(o.MODE_DENSITY).cpu().numpy()[..., 0].astype(np.float64)
        st = eng.stats()
    finally:
        eng.close()
    ref = ref[..., 0]
    assert st["direct_vector_reds"] == 0 and st["particles_huge"] > 0, st       # nothing was splatted by K1 itself
    big = ref > FLOOR * ref.max()
    assert (np.abs(img[big] - ref[big]) / ref[big]).max() <= REL_TOL
    assert np.abs(img[~big] - ref[~big]).max() <= 2 * FLOOR * ref.max()


def test_pair_capacity_overflow_falls_back_to_atomics(oracle_lut):
    """250k footprints that each cover all 32 gather tiles of a 256^2 image: 8M (record, tile) pairs against a capacity of
    6 per queue slot (striped over 32 reservation counters) -> part of the records must take the cooperative atomic path,
    and the image must not change."""
    from topsy_b200.engine import SplatEngine
    rs = np.random.RandomState(3)
    n, R = 250_000, 256
    pos = rs.uniform(-0.2, 0.2, (n, 3)).astype(np.float32)
    h = rs.uniform(0.8, 1.6, n).astype(np.float32)            # quad width 2 h R / scale = 410 .. 820 px
    m = rs.uniform(0.5, 1.5, n).astype(np.float32)
    M = o.transform_matrix(o.rotate(np.eye(3), 0.2, -0.3), np.zeros(3), 1.0); sf = o.scale_factor(1.0)
    ref = co.splat(pos[:, 0], pos[:, 1], pos[:, 2], h, (m,), M, sf, R, o.MODE_DENSITY, oracle_lut)
    eng = SplatEngine(R)
    try:
        eng.set_kernel_lut(oracle_lut)
        eng.set_camera(M, sf)
        x, y, z, hd, md = _to_dev(pos[:, 0], pos[:, 1], pos[:, 2], h, m)
        eng.set_particles(x, y, z, hd); eng.set_weights(md)
        img = eng.render(o.MODE_DENSITY).cpu().numpy()
        st = eng.stats()
        assert st["particles_huge"] > 0 and st["particles_tiled"] > 0, st      # both routes were taken
        assert_image_parity(img, ref, "pair overflow")
    finally:
        eng.close()


def test_oversize_range_in_multi_range_call_is_split(oracle_lut):
    """a multi-range call in which ONE range exceeds the scratch queue capacity (a clustered
    snapshot with one huge cell) used to fail with 'a single range exceeds the scratch capacity'; it must be split
    like the single-range path does.  Scratch is sized for 2^20 particles per call, the first range holds 2.5 M."""
    from topsy_b200.engine import SplatEngine
    rs = np.random.RandomState(17)
    n, R = 3_000_000, 256
    pos = rs.uniform(-1, 1, (n, 3)).astype(np.float32)
    h = (0.004 * np.exp(rs.normal(size=n) * 0.6)).astype(np.float32)        # ~0.5 px .. 3 px: direct + some deferred
    h[::1000] *= 30.0                                                        # a few big footprints in every range
    m = rs.uniform(0.5, 1.5, n).astype(np.float32)
    M = o.transform_matrix(o.rotate(np.eye(3), 0.1, 0.2), np.zeros(3), 1.0); sf = o.scale_factor(1.0)
    starts = np.array([3, 2_500_001, 2_800_000], np.int64)
    lens = np.array([2_499_998, 250_000, 199_999], np.int64)
    ref = co.splat(pos[:, 0], pos[:, 1], pos[:, 2], h, (m,), M, sf, R, o.MODE_DENSITY, oracle_lut, ranges=(starts, lens))
    eng = SplatEngine(R, max_particles_per_call=1 << 20)
    try:
        eng.set_kernel_lut(oracle_lut)
        eng.set_camera(M, sf)
        x, y, z, hd, md = _to_dev(pos[:, 0], pos[:, 1], pos[:, 2], h, m)
        eng.set_particles(x, y, z, hd); eng.set_weights(md)
        img = eng.render(o.MODE_DENSITY, starts, lens).cpu().numpy()
        assert eng.stats()["particles_submitted"] == lens.sum()
        assert_image_parity(img, ref, "oversize range")
    finally:
        eng.close()


def test_more_ranges_than_one_table_holds(engine200, oracle_lut):
    """more than 65536 ranges in one call (ArrayDataLoader with nside >= 41) are processed in batches."""
    rs = np.random.RandomState(23)
    n = 70_000 * 3
    pos = rs.uniform(-1, 1, (n, 3)).astype(np.float32)
    h = (0.01 * np.exp(rs.normal(size=n) * 0.5)).astype(np.float32)
    m = rs.uniform(0.5, 1.5, n).astype(np.float32)
    M = o.transform_matrix(np.eye(3), np.zeros(3), 1.0); sf = o.scale_factor(1.0)
    starts = np.arange(70_000, dtype=np.int64) * 3
    lens = np.full(70_000, 2, np.int64)
    ref = co.splat(pos[:, 0], pos[:, 1], pos[:, 2], h, (m,), M, sf, 200, o.MODE_DENSITY, oracle_lut, ranges=(starts, lens))
    x, y, z, hd, md = _to_dev(pos[:, 0], pos[:, 1], pos[:, 2], h, m)
    engine200.set_kernel_lut(oracle_lut)
    engine200.set_camera(M, sf)
    engine200.set_particles(x, y, z, hd); engine200.set_weights(md)
    img = engine200.render(o.MODE_DENSITY, starts, lens).cpu().numpy()
    assert engine200.stats()["particles_submitted"] == lens.sum()
    assert_image_parity(img, ref, "70000 ranges")
