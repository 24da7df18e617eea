"""SplitBuffers address arithmetic (no GPU)."""
import numpy as np
import pytest

from topsy_b200.split_buffers import SplitBuffers


@pytest.fixture
def sb():
    return SplitBuffers(50, 15)        # buffers of 15, 15, 15, 5


def test_layout(sb):
    assert sb.num_buffers == 4
    assert list(sb._buffer_particle_sizes) == [15, 15, 15, 5]
    assert SplitBuffers(10, 15).num_buffers == 1
    assert SplitBuffers(30, 15).num_buffers == 2


def test_global_to_split(sb):
    cases = {(0, 10): ([0], [0], [10]), (0, 20): ([0, 1], [0, 0], [15, 5]), (0, 45): ([0, 1, 2], [0, 0, 0], [15, 15, 15]),
             (15, 10): ([1], [0], [10]), (14, 2): ([0, 1], [14, 0], [1, 1]), (20, 20): ([1, 2], [5, 0], [10, 10]),
             (49, 1): ([3], [4], [1]), (0, 50): ([0, 1, 2, 3], [0, 0, 0, 0], [15, 15, 15, 5])}
    for (start, length), want in cases.items():
        assert sb.global_to_split(start, length) == want
    with pytest.raises(ValueError):
        sb.global_to_split(0, 100)
    with pytest.raises(ValueError):
        sb.global_to_split(49, 2)


def test_monotonic_sweep_equals_per_range_search(sb):
    rs = np.random.RandomState(1337)
    for _ in range(200):
        cuts = np.sort(rs.randint(0, 50, size=6))
        starts = cuts[:-1]
        lens = rs.randint(np.diff(cuts) + 1)
        keep = lens != 0
        starts, lens = starts[keep], lens[keep]
        fast = sb.global_to_split_monotonic(starts, lens)
        slow = [([], []) for _ in range(sb.num_buffers)]
        for s, l in zip(starts, lens):
            for b, ls, ll in zip(*sb.global_to_split(s, l)):
                slow[b][0].append(ls)
                slow[b][1].append(ll)
        assert fast == slow
    with pytest.raises(ValueError):
        sb.global_to_split_monotonic([45], [10])
