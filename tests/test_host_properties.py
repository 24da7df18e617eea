"""Property tests (hypothesis) of the host-side scheduling arithmetic: split-buffer mapping and per-cell range mapping.
Size-independent invariants of rows a7 / a8 of SURVEY.md section 8; no GPU."""
import numpy as np
from hypothesis import given, settings, strategies as st

from topsy_b200.cell_layout import CellLayout
from topsy_b200.progressive_render import RenderProgressionWithCells
from topsy_b200.split_buffers import SplitBuffers


@st.composite
def monotone_ranges(draw):
    n = draw(st.integers(1, 5000))
    per_buffer = draw(st.integers(1, 2000))
    cuts = sorted(draw(st.lists(st.integers(0, n), min_size=2, max_size=40)))
    starts, lens = [], []
    for a, b in zip(cuts[:-1], cuts[1:]):
        starts.append(a)
        lens.append(draw(st.integers(0, b - a)))
    return n, per_buffer, np.array(starts, np.int64), np.array(lens, np.int64)


@settings(max_examples=200, deadline=None)
@given(monotone_ranges())
def test_split_mapping_is_a_partition_of_the_requested_particles(case):
    n, per_buffer, starts, lens = case
    sb = SplitBuffers(n, per_buffer)
    pieces = sb.global_to_split_monotonic(starts, lens)
    assert len(pieces) == sb.num_buffers
    want = np.zeros(n, np.int32)
    for s, l in zip(starts, lens):
        want[s:s + l] += 1
    got = np.zeros(n, np.int32)
    for k, (ls, ll) in enumerate(pieces):
        b0, b1 = sb.buffer_range(k)
        assert len(ls) == len(ll)
        for s, l in zip(ls, ll):
            assert l > 0 and 0 <= s and s + l <= b1 - b0         # inside the buffer, never empty
            got[b0 + s:b0 + s + l] += 1
    assert np.array_equal(got, want)


@settings(max_examples=60, deadline=None)
@given(st.integers(200, 4000), st.integers(2, 6), st.integers(0, 2 ** 31 - 1), st.integers(2, 9))
def test_cell_mapped_blocks_tile_every_cell_exactly_once(n, nside, seed, n_blocks):
    """Consecutive logical blocks [0, b1), [b1, b2), ... mapped through the cells cover every particle of every selected
    cell exactly once between them (progressive_render.py:152-187): a complete sequence of REFINE frames renders the same
    particles as one EXPORT frame, whatever the block sizes."""
    rs = np.random.RandomState(seed)
    pos = rs.uniform(-1, 1, (n, 3))
    layout, _ = CellLayout.from_positions(pos, -1.0001, 1.0001, nside)
    rp = RenderProgressionWithCells(layout, n, 100)
    cuts = np.unique(np.concatenate([[0, n], rs.randint(0, n + 1, n_blocks - 1)]))
    hits = np.zeros(n, np.int32)
    for a, b in zip(cuts[:-1], cuts[1:]):
        s, l = rp._map_logical_range_to_actual_ranges(int(a), int(b - a))
        assert (np.asarray(l) > 0).all()
        for si, li in zip(s, l):
            hits[si:si + li] += 1
    assert (hits == 1).all()
