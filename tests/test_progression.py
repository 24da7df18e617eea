"""Host scheduling layer: RenderProgression / RenderProgressionWithCells (no GPU).

Same behaviours the reference pins in its tests/test_progression.py, plus differential checks against outputs of the
reference's own module (tests/golden/scheduling_goldens.npz, produced by tests/golden/make_golden.py)."""
import numpy as np
import pytest

from topsy_b200 import config, progressive_render
from topsy_b200.cell_layout import CellLayout
from topsy_b200.drawreason import DrawReason
from topsy_b200.progressive_render import RenderProgression, RenderProgressionWithCells

FRAME = 1.0 / config.TARGET_FPS


def single(block):
    starts, lens = block
    assert len(starts) == 1 and len(lens) == 1
    return starts[0], lens[0]


def test_first_block_is_capped_by_the_initial_budget():
    n0 = int(config.INITIAL_PARTICLES_TO_RENDER)
    small = RenderProgression(n0 // 2)
    small.start_frame(DrawReason.INITIAL_UPDATE)
    assert single(small.get_block(0.0)) == (0, n0 // 2)
    large = RenderProgression(n0 * 2)
    large.start_frame(DrawReason.INITIAL_UPDATE)
    assert single(large.get_block(0.0)) == (0, n0)


def test_export_renders_everything_in_one_block_when_small():
    rp = RenderProgression(int(config.INITIAL_PARTICLES_TO_RENDER) * 2)
    assert rp.start_frame(DrawReason.EXPORT) is True
    assert single(rp.get_block(0.0)) == (0, int(config.INITIAL_PARTICLES_TO_RENDER) * 2)
    rp.end_block(0.1)
    assert rp.get_block(1.0) is None
    rp2 = RenderProgression(1000, 100)
    rp2.start_frame(DrawReason.EXPORT)
    assert single(rp2.get_block(0.0)) == (0, 1000)


def test_export_is_chunked_for_huge_snapshots():
    chunk = config.MAX_PARTICLES_PER_EXPORT_RENDERCALL
    rp = RenderProgression(chunk * 5)
    rp.start_frame(DrawReason.EXPORT)
    for k in range(5):
        start, count = single(rp.get_block(100.0 * k))      # elapsed time is irrelevant in EXPORT frames
        assert (start, count) == (chunk * k, chunk)
        rp.end_block(100.0 * (k + 1))
    assert rp.get_block(500.0) is None
    assert rp.start_frame(DrawReason.EXPORT) is True


def test_interactive_frame_uses_remaining_time():
    rp = RenderProgression(1000, 100)
    rp.start_frame(DrawReason.CHANGE)
    assert single(rp.get_block(0.0)) == (0, 100)
    rp.end_block(0.5 * FRAME)
    assert single(rp.get_block(0.5 * FRAME)) == (100, 50)      # half the frame left -> half the budget
    rp.end_block(FRAME)
    assert rp.get_block(FRAME) is None
    assert rp.end_frame_get_scalefactor() == 1000.0 / 150


def test_slow_frame_then_refine():
    rp = RenderProgression(1000, 100)
    rp.start_frame(DrawReason.CHANGE)
    assert rp.get_block(0.0) is not None
    rp.end_block(1.0)                                           # 30x over budget
    assert rp.get_block(1.0) is None
    assert rp.end_frame_get_scalefactor() == 10.0
    assert rp.needs_refine()
    assert rp.start_frame(DrawReason.REFINE) is False           # keep the image
    assert single(rp.get_block(0.0)) == (100, int(100 / config.TARGET_FPS))


def test_at_least_one_block_and_one_particle():
    rp = RenderProgression(1000, 100)
    rp.start_frame(DrawReason.CHANGE)
    assert rp.get_block(1.0) is not None                        # first block is unconditional
    rp = RenderProgression(1000, 3)
    rp.start_frame(DrawReason.CHANGE)
    rp.get_block(0.0)
    rp.end_block(1.0)
    assert rp.get_block(1.0) is None
    rp.end_frame_get_scalefactor()
    rp.start_frame(DrawReason.REFINE)
    assert single(rp.get_block(1.0)) == (3, 1)                  # never recommends zero particles


def test_presentation_change_renders_nothing():
    rp = RenderProgression(1000, 100)
    rp.start_frame(DrawReason.CHANGE)
    t = 0.0
    while rp.get_block(t) is not None:
        t += 1e-5
        rp.end_block(t)
    rp.end_frame_get_scalefactor()
    assert not rp.needs_refine()
    rp.start_frame(DrawReason.PRESENTATION_CHANGE)
    assert rp.get_block(0.0) is None
    rp.end_frame_get_scalefactor()
    assert not rp.needs_refine()


def test_get_block_needs_a_frame():
    with pytest.raises(RuntimeError):
        RenderProgression(1000, 100).get_block(0.0)


def test_adaptive_budget_matches_reference_trace(sched_goldens):
    trace = sched_goldens["rpp_trace"]
    rp = RenderProgression(10 ** 7)
    for (start, count, sf, rec), frame_time in zip(trace, [0.01, 0.2, 0.05, 0.033, 0.001, 0.5]):
        rp.start_frame(DrawReason.CHANGE)
        blk = rp.get_block(0.0)
        assert single(blk) == (int(start), int(count))
        rp.end_block(frame_time)
        assert rp.end_frame_get_scalefactor() == sf
        assert rp._recommended_num_particles_to_render == int(rec)


@pytest.fixture
def cells_and_positions():
    rs = np.random.RandomState(1337)
    pos = rs.uniform(0.0, 1.0, (100000, 3))
    layout, order = CellLayout.from_positions(pos, 0.0, 1.0, 10)
    return RenderProgressionWithCells(layout, len(pos), 100), pos[order]


def test_cell_blocks_cover_every_particle_once(cells_and_positions):
    rp, pos = cells_and_positions
    layout = rp._cell_layout
    hits = np.zeros(len(pos), dtype=np.int32)
    rp.start_frame(DrawReason.CHANGE)
    first = True
    while True:
        starts, lens = rp.get_block(0.0)
        for s, l in zip(starts, lens):
            assert l > 0
            assert layout.cell_index_from_offset(s) == layout.cell_index_from_offset(s + l - 1)   # never straddles
            hits[s:s + l] += 1
        if first:
            assert 95 < hits.sum() < 105
            first = False
        rp.end_block(0.0001)
        rp.end_frame_get_scalefactor()
        if not rp.needs_refine():
            break
        rp.start_frame(DrawReason.REFINE)
    assert (hits == 1).all()
    rp.start_frame(DrawReason.CHANGE)
    total = 0
    while (blk := rp.get_block(0.0)):
        total += int(np.sum(blk[1]))
        rp.end_block(0.0)
    assert total == len(pos)


def test_sphere_selection_limits_the_blocks(cells_and_positions):
    rp, pos = cells_and_positions
    rp.select_sphere((0.5, 0.5, 0.5), 0.1)
    rp.start_frame(DrawReason.CHANGE)
    hits = np.zeros(len(pos), dtype=np.int32)
    while (blk := rp.get_block(0.0)):
        for s, l in zip(*blk):
            hits[s:s + l] += 1
        rp.end_block(0.0)
    assert hits.max() == 1
    r = np.linalg.norm(pos - 0.5, axis=1)
    assert (r[hits == 1] < 0.4).all()
    assert (r[hits == 0] > 0.1).all()
    assert rp.get_fraction_volume_selected() < 0.2


def test_cell_mapping_matches_reference_module(sched_goldens):
    g = sched_goldens
    layout, order = CellLayout.from_positions(g["cl64_pos"], -1.0, 1.0, 10)
    rp = RenderProgressionWithCells(layout, len(g["cl64_pos"]), 100)
    assert np.array_equal(rp._cell_phase_shifts, g["rp_phase"])
    rp.start_frame(DrawReason.CHANGE)
    for i in range(int(g["rp_nblocks"])):
        blk = rp.get_block(0.0)
        assert np.array_equal(np.stack([np.asarray(blk[0]), np.asarray(blk[1])]), g[f"rp_block{i}"]), i
        rp.end_block(0.0001)
        rp.end_frame_get_scalefactor()
        if rp.needs_refine():
            rp.start_frame(DrawReason.REFINE)
    rp2 = RenderProgressionWithCells(layout, len(g["cl64_pos"]), 100)
    rp2.select_sphere((0.1, -0.2, 0.3), 0.35)
    assert rp2.get_fraction_volume_selected() == float(g["rp_sphere_fraction"])
    rp2.start_frame(DrawReason.EXPORT)
    blk = rp2.get_block(0.0)
    assert np.array_equal(np.stack([np.asarray(blk[0]), np.asarray(blk[1])]), g["rp_sphere_export_block"])
    s, l = rp2._map_logical_range_to_actual_ranges(1234, 4321)
    assert np.array_equal(np.stack([s, l]), g["rp_sphere_map_1234_4321"])
