"""Host-side task planning (no GPU needed): the invariants the kernels rely on, checked on random ragged walks
through the test hook pb_debug_plan.

* every (32-wide i-block, list entry) of every walk is covered by exactly one task of the right kind;
* chunk starts are multiples of the 256-entry j tile (hence of 8) and list offsets multiples of 4 entries: every index tile the force kernel
  fetches with a bulk (TMA) copy then starts on a 16-byte boundary;
* nib * jsplit <= 8 warps, i_first a multiple of 32, groups follow the binary 8/4/2/1 decomposition;
* partial-sum slots of different (task, i-block) pairs never overlap and the reduction tables point at them;
* fused reduction: a task names its group's first i-block (`blk0`) and the number of partial sums each of the group's blocks
  will receive (`n_chunks`) — the counter target of the warp that decides whether it delivered the last one."""
import numpy as np
import pytest

from petar_b200 import engine


def _check(n_epi, n_epj, n_spj, n_streams):
    walks, tasks, ibl, n_part = engine.debug_plan(n_epi, n_epj, n_spj, n_streams)
    nw = len(n_epi)
    assert np.array_equal(walks[:, 1], n_epi) and np.array_equal(walks[:, 3], n_epj) and np.array_equal(walks[:, 5], n_spj)
    assert np.array_equal(walks[:, 0], np.concatenate([[0], np.cumsum(n_epi)[:-1]]))
    assert np.all(walks[:, 2] % 4 == 0) and np.all(walks[:, 4] % 4 == 0)               # 16-byte aligned lists
    # lists do not overlap
    for col, cnt in ((2, 3), (4, 5)):
        order = np.argsort(walks[:, col], kind="stable")
        ends = walks[order, col] + walks[order, cnt]
        assert np.all(ends[:-1] <= walks[order, col][1:])
    used = np.zeros(n_part, dtype=np.int32)
    cover = {}
    delivered = np.zeros(len(ibl), dtype=np.int32)
    for walk, i_first, nib, jsplit, kind, j_begin, j_count, part_base, blk0, n_chunks in tasks:
        # fused reduction: this task delivers one partial sum to each of blocks blk0 .. blk0 + nib - 1, at a slot the
        # block's reduction record reaches with one of its n_chunks strides
        for b in range(nib):
            bp, bn, bs, bo, bv = ibl[blk0 + b]
            assert bn == n_chunks and bs == nib * 32
            k, r = divmod(part_base + b * 32 - bp, bs)
            assert r == 0 and 0 <= k < bn
            delivered[blk0 + b] += 1
        assert 0 <= walk < nw and kind in (0, 1) and nib in (1, 2, 4, 8) and nib * jsplit == 8
        assert i_first % 32 == 0 and i_first < max(1, n_epi[walk]) and j_begin % 8 == 0 and j_count > 0
        nj = n_epj[walk] if kind == 0 else n_spj[walk]
        assert j_begin + j_count <= nj
        assert part_base >= 0 and part_base + nib * 32 <= n_part
        used[part_base:part_base + nib * 32] += 1
        for b in range(nib):
            cover.setdefault((walk, i_first // 32 + b, kind), []).append((j_begin, j_count))
    assert used.max(initial=0) <= 1                                                    # partial slots never shared
    for w in range(nw):
        for b in range((n_epi[w] + 31) // 32):
            for kind, nj in ((0, n_epj[w]), (1, n_spj[w])):
                segs = sorted(cover.get((w, b, kind), []))
                pos = 0
                for jb, jc in segs:                                                    # contiguous, no gaps, no overlaps
                    assert jb == pos
                    pos += jc
                assert pos == nj
    # reduction tables: every i-block once, chunks = tasks that cover it, slots inside the partial array
    assert len(ibl) == sum((n + 31) // 32 for n in n_epi)
    assert np.array_equal(delivered, ibl[:, 1]) if len(ibl) else True              # every block gets exactly n_chunks partial sums
    out_seen = set()
    for part_base, n_chunks, stride, out_off, n_valid in ibl:
        assert 1 <= n_valid <= 32 and out_off not in out_seen and (n_chunks == 0 or part_base + (n_chunks - 1) * stride + 32 <= n_part)
        out_seen.add(out_off)
    return len(tasks)


@pytest.mark.parametrize("seed", range(6))
def test_random_ragged_walks(seed):
    rng = np.random.default_rng(seed)
    nw = int(rng.integers(1, 60))
    n_epi = rng.integers(1, 700, nw).astype(np.int32)
    n_epj = rng.integers(0, 9000, nw).astype(np.int32)
    n_spj = rng.integers(0, 12000, nw).astype(np.int32)
    n_epj[rng.random(nw) < 0.15] = 0
    n_spj[rng.random(nw) < 0.15] = 0
    _check(n_epi, n_epj, n_spj, int(rng.integers(1, 9)))


def test_edge_shapes():
    _check(np.array([1], np.int32), np.array([1], np.int32), np.array([0], np.int32), 1)           # one pair
    _check(np.array([33], np.int32), np.array([7], np.int32), np.array([9], np.int32), 4)          # 2 i-blocks, tiny lists
    _check(np.array([512], np.int32), np.array([200000], np.int32), np.array([150000], np.int32), 1)  # long lists, many chunks
    _check(np.array([5, 64, 255, 256, 257], np.int32), np.array([0, 1, 8, 9, 4096], np.int32), np.array([3, 0, 0, 17, 4097], np.int32), 8)
    assert _check(np.array([40], np.int32), np.array([0], np.int32), np.array([0], np.int32), 1) == 0   # nothing to do
