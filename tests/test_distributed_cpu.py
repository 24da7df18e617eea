"""Multi-GPU host logic on CPU: sharding arithmetic, and the image sum over a world_size-2 gloo group."""
import os
import socket

import numpy as np
import pytest

from topsy_b200 import distributed as D
from topsy_b200.cell_layout import CellLayout
from topsy_b200.drawreason import DrawReason
from topsy_b200.progressive_render import RenderProgressionWithCells


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_striping_partitions_every_cell(world):
    rs = np.random.RandomState(1)
    lengths = rs.randint(0, 50, 64)
    lengths[:3] = [0, 1, 7]
    offsets = np.cumsum(lengths) - lengths
    seen = np.zeros(lengths.sum(), dtype=int)
    for r in range(world):
        idx = D.shard_indices(offsets, lengths, r, world)
        mine = D.shard_cell_lengths(lengths, r, world)
        assert len(idx) == mine.sum()
        assert (np.diff(idx) > 0).all() if len(idx) > 1 else True     # monotonic: cells stay contiguous and ordered
        seen[idx] += 1
        assert (mine <= -(-lengths // world)).all() and (mine >= lengths // world).all()     # balanced inside each cell
    assert (seen == 1).all()
    assert sum(D.shard_cell_lengths(lengths, r, world) for r in range(world)).tolist() == lengths.tolist()


def test_shards_keep_cell_structure_for_the_progression():
    """Each rank can run the unchanged progression on its own per-cell lengths: blocks never straddle cells and together
    cover the shard exactly once."""
    rs = np.random.RandomState(4)
    pos = rs.uniform(0, 1, (20000, 3))
    layout, order = CellLayout.from_positions(pos, 0.0, 1.0, 8)
    for rank in range(2):
        lens = D.shard_cell_lengths(layout._lengths, rank, 2)
        offs = np.cumsum(lens) - lens
        local = CellLayout(layout._centres, offs, lens)
        rp = RenderProgressionWithCells(local, int(lens.sum()), 500)
        hits = np.zeros(int(lens.sum()), dtype=int)
        rp.start_frame(DrawReason.CHANGE)
        while True:
            for s, l in zip(*rp.get_block(0.0)):
                hits[s:s + l] += 1
            rp.end_block(1.0)
            rp.end_frame_get_scalefactor()
            if not rp.needs_refine():
                break
            rp.start_frame(DrawReason.REFINE)
        assert (hits == 1).all()


def test_presentation_slabs_tile_the_image_and_balance_the_link_cost():
    """The slabs of the fused reduce + colormap kernel: contiguous, cover every row once, and sized so that the gather
    destination (which also receives everybody's RGBA rows) is busy no longer than the other ranks."""
    k = D.REMOTE_STORE_COST
    for R in (200, 2048, 4097):
        for world in (1, 2, 3, 8):
            for in_b, out_b in ((4, 4), (8, 4), (16, 4), (16, 16), (4, 16)):
                slabs = [D.presentation_slab(R, r, world, in_b, out_b) for r in range(world)]
                assert slabs[0][0] == 0 and sum(n for _, n in slabs) == R
                for (a, n), (b, _) in zip(slabs, slabs[1:]):
                    assert a + n == b
                if world == 1:
                    continue
                rows = [n for _, n in slabs]

                def slowest(rows_dst, rows_other):      # link work of the busiest rank, in units of loaded bytes
                    return max((world - 1) * rows_dst * R * in_b + k * (R - rows_dst) * R * out_b,
                               rows_other * R * ((world - 1) * in_b + k * out_b))

                got = slowest(rows[0], max(rows[1:]))
                assert got <= slowest(-(-R // world), -(-R // world)) * (1 + 1e-12)          # never worse than equal slabs
                for n_dst in range(0, R + 1, max(1, R // 64)):                               # ... or than any other split
                    assert got <= slowest(n_dst, -(-(R - n_dst) // (world - 1))) * (1 + 1e-12)
    assert abs(D.presentation_slab(4096, 0, 2, 4, 4)[1] - 2048) <= 1     # two GPUs: equal slabs
    assert 250 < D.presentation_slab(4096, 0, 8, 4, 4)[1] < 350          # c5 on 8 GPUs: ~7 % of the rows instead of 12.5 %


def test_row_slabs_tile_the_image():
    for R in (200, 2048, 4097):
        for world in (1, 2, 3, 8):
            rows = [D.row_slab(R, r, world) for r in range(world)]
            assert rows[0][0] == 0 and sum(n for _, n in rows) == R
            for (a, n), (b, _) in zip(rows, rows[1:]):
                assert a + n == b


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rs = np.random.RandomState(100 + rank)
        part = torch.from_numpy(rs.uniform(size=(32, 32, 2)).astype(np.float32))
        total = D.reduce_image_host(part.clone())
        q.put((rank, total.numpy(), part.numpy()))
        only0 = D.reduce_image_host(part.clone(), dst=0)
        if rank == 0:
            assert torch.equal(only0, total)
    finally:
        dist.destroy_process_group()


def test_image_sum_over_gloo_world_of_two():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    got.sort(key=lambda t: t[0])
    want = got[0][2] + got[1][2]
    np.testing.assert_allclose(got[0][1], want, rtol=1e-6)
    np.testing.assert_allclose(got[1][1], want, rtol=1e-6)


@pytest.mark.parametrize("world", [2, 3])
def test_shard_loader_stripes_partition_the_snapshot(world):
    """distributed.shard_loader: every rank keeps its per-cell stripe; the stripes partition the snapshot and each
    rank's loader (len, getters, cell layout, progression) describes exactly its stripe."""
    from topsy_b200 import loader
    rs = np.random.RandomState(8)
    n = 5000
    pos = rs.uniform(-1, 1, (n, 3)).astype(np.float32)
    smooth = rs.uniform(0.01, 0.1, n).astype(np.float32)
    mass = np.arange(n, dtype=np.float32)                      # a unique tag per particle
    q = rs.normal(size=n).astype(np.float32)
    np.random.seed(3)
    seen = np.zeros(n, int)
    for rank in range(world):
        np.random.seed(3)                                        # same within-cell shuffle on every rank
        ld = loader.ArrayDataLoader(None, pos, smooth, mass, quantities={"q": q}, nside=4)
        full_lengths = ld._cell_layout._lengths.copy()
        D.shard_loader(ld, rank, world)
        assert ld.global_num_particles == n
        assert len(ld) == D.shard_cell_lengths(full_lengths, rank, world).sum()
        tags = ld.get_mass().astype(int)
        seen[tags] += 1
        np.testing.assert_array_equal(ld.get_positions(), pos[tags])
        np.testing.assert_array_equal(ld.get_named_quantity("q"), q[tags])
        np.testing.assert_array_equal(ld._cell_layout._lengths, D.shard_cell_lengths(full_lengths, rank, world))
        assert ld.get_render_progression()._cell_layout.get_num_particles() == len(ld)
        # particles of the stripe are still grouped by cell
        sl = ld._cell_layout.cell_slice(int(np.argmax(ld._cell_layout._lengths)))
        cell_pos = ld.get_positions()[sl]
        assert np.ptp(cell_pos, axis=0).max() <= 2.0 * 1.05 / 4 + 1e-6
    assert (seen == 1).all()
    # a loader without cells: every world-th particle
    t0 = loader.TestDataLoader(None, 1000)
    t1 = loader.TestDataLoader(None, 1000)
    D.shard_loader(t1, 1, world)
    np.testing.assert_array_equal(t1.get_positions(), t0.get_positions()[1::world])
    np.testing.assert_array_equal(t1.get_smooth(), t0.get_smooth()[1::world])
    assert len(t1.get_mass()) == len(t1) == len(range(1, 1000, world))


@pytest.mark.parametrize("world,n_total", [(2, 100_003), (3, 50_000), (8, 9_000)])
def test_synthetic_stripes_are_slices_of_one_snapshot(world, n_total):
    """bench.py's N > 1 parity check rests on this: rank r of G generates exactly the particles that per-cell striping
    (shard_indices) picks from the single-GPU snapshot, bit for bit, and that snapshot is already in topsy's cell order.
    Also covers stripes with empty cells (9000 particles over 4096 cells and 8 ranks)."""
    torch = pytest.importorskip("torch")
    from topsy_b200 import synthetic
    wl = synthetic.WORKLOADS["c3"]                                       # weighted: x y z h m q
    full, lengths = synthetic.generate_striped(wl, "cpu", n_total=n_total, chunk=7001)
    assert int(lengths.sum()) == n_total
    offsets = np.cumsum(lengths.numpy()) - lengths.numpy()
    seen = 0
    for rank in range(world):
        part, mine = synthetic.generate_striped(wl, "cpu", n_total=n_total, rank=rank, world=world, chunk=4999)
        idx = D.shard_indices(offsets, lengths.numpy(), rank, world)
        assert np.array_equal(mine.numpy(), D.shard_cell_lengths(lengths.numpy(), rank, world))
        assert synthetic.stripe_size(n_total, rank, world) == len(idx)
        for name, arr in part.items():
            assert torch.equal(arr, full[name][torch.from_numpy(idx)]), (rank, name)
        seen += len(idx)
    assert seen == n_total
    # the snapshot is in cell order: the reference's cell layout of it is the identity permutation
    pos = np.stack([full[k].numpy() for k in "xyz"], axis=1)
    layout, ordering = CellLayout.from_positions(pos, -0.5 * synthetic.BOX, 0.5 * synthetic.BOX, synthetic.NSIDE)
    assert np.array_equal(ordering, np.arange(n_total))
    assert np.array_equal(layout._lengths, lengths.numpy())
