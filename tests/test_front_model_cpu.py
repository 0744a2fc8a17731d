"""The thin-front flood solver's scheme (csrc/flood.cu, flood_front_kernel), stated as a small host model and checked against the oracle's key field.

What the kernel relies on, and what this file checks without a GPU: label-correcting relaxation of keys (dist << 15 | order) with lists of
(cell, key) pairs — expand a pair = lower every neighbour that is not a wall and lies above key + 1 level, list every cell lowered with the key
written — reaches the oracle's key field (FloodFracturer.cpp:98-191 under the lowest-seed-index rule, oracle flood_keys)

  * whatever order the pairs of a step are expanded in and however the pairs are dealt to the CTAs (own list / another CTA's inbox),
  * however many steps a CTA runs on its own list between two barriers (fronts of different CTAs run ahead of each other and correct each other),
  * with out-of-date pairs in the lists (a cell lowered twice is listed twice), and
  * when pairs are LOST (a list overflowed): the keys are then upper bounds realised by real paths, and a plain relaxation to the fixed point from
    that state — what the tile worklist does after the hand-over — ends at the same field.

The GPU tests (tests/test_flood_gpu.py::test_result_does_not_depend_on_the_front_limit) check the kernel itself, bit for bit.
"""
import numpy as np
import pytest

from conftest import pick_seeds, random_blob_grid

WALL, UNREACHED, LEVEL = 0xFFFFFFFF, 0xFFFFFFFE, 1 << 15


def _offsets(nneigh):
    if nneigh == 6:
        return [(1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1)]
    return [(dx, dy, dz) for dx in (-1, 0, 1) for dy in (-1, 0, 1) for dz in (-1, 0, 1) if (dx, dy, dz) != (0, 0, 0)]


def _relax_to_fixed_point(keys, offs):
    """what the tiles do after a hand-over: key(v) = min(key(v), min over neighbours + 1 level) until nothing changes"""
    X, Y, Z = keys.shape
    changed = True
    while changed:
        changed = False
        for order in (1, -1):
            cells = list(np.ndindex(X, Y, Z))[::order]
            for x, y, z in cells:
                k = int(keys[x, y, z])
                if k == WALL:
                    continue
                best = k
                for dx, dy, dz in offs:
                    a, b, c = x + dx, y + dy, z + dz
                    if 0 <= a < X and 0 <= b < Y and 0 <= c < Z:
                        m = int(keys[a, b, c])
                        if m < UNREACHED - LEVEL and m + LEVEL < best:
                            best = m + LEVEL
                if best < k:
                    keys[x, y, z] = best
                    changed = True
    return keys


def front_model(grid, seeds, nneigh, rng, ctas=4, sublevels=3, far_prob=0.3, lose_prob=0.0, limit=1 << 30):
    """-> (keys, handed_over).  Mirrors flood_front_seed_kernel + flood_front_kernel: seeds planted in order, the first front dealt round-robin,
    `sublevels` steps per CTA between two barriers, claims pushed to the own list or dealt into another CTA's inbox, inboxes emptied at the barrier."""
    X, Y, Z = grid.shape
    offs = _offsets(nneigh)
    keys = np.where(grid == 0, WALL, UNREACHED).astype(np.uint64)
    for s, (x, y, z, _) in enumerate(seeds):
        keys[x, y, z] = s  # a later seed on the same cell overwrites; a seed on an EMPTY cell is a source all the same
    first = [(int(x), int(y), int(z)) for s, (x, y, z, _) in enumerate(seeds) if keys[x, y, z] == s]
    lists = [[] for _ in range(ctas)]
    for i, cell in enumerate(first):
        lists[i % ctas].append((cell, int(keys[cell])))
    lost = False
    while any(lists):
        if sum(len(l) for l in lists) > limit or lost:
            return _relax_to_fixed_point(keys, offs), True
        inbox = [[] for _ in range(ctas)]
        for cta in rng.permutation(ctas):  # the CTAs of an interval in any order: each sees whatever the others have written so far
            cur = lists[cta]
            for _ in range(sublevels):
                nxt = []
                for j in rng.permutation(len(cur)):
                    (x, y, z), own = cur[j]
                    nk = own + LEVEL
                    for dx, dy, dz in offs:
                        a, b, c = x + dx, y + dy, z + dz
                        if not (0 <= a < X and 0 <= b < Y and 0 <= c < Z):
                            continue
                        old = int(keys[a, b, c])
                        if old != WALL and old > nk:  # atomicMin lowered it: unreached, or a correction
                            keys[a, b, c] = nk
                            if rng.rand() < lose_prob:
                                lost = True  # the list was full: the cell is lowered but not listed
                            elif rng.rand() < far_prob:
                                inbox[rng.randint(ctas)].append(((a, b, c), nk))
                            else:
                                nxt.append(((a, b, c), nk))
                cur = nxt
            lists[cta] = cur
        for cta in range(ctas):
            lists[cta] = lists[cta] + inbox[cta]
    if lost:
        return _relax_to_fixed_point(keys, offs), True
    return keys, False


CASES = [
    dict(ctas=1, sublevels=1, far_prob=0.0),                 # a plain level-synchronous BFS
    dict(ctas=4, sublevels=1, far_prob=0.5),
    dict(ctas=4, sublevels=5, far_prob=0.0),                 # fronts run ahead of each other by up to 5 levels
    dict(ctas=3, sublevels=16, far_prob=0.3),
    dict(ctas=4, sublevels=4, far_prob=0.3, lose_prob=0.02),  # pairs are lost: the relaxation that follows must repair everything
    dict(ctas=4, sublevels=4, far_prob=0.3, limit=40),        # hand-over once more than 40 pairs are pending
]


@pytest.mark.parametrize("case", range(len(CASES)))
@pytest.mark.parametrize("dfunc,nneigh", [(1, 6), (2, 26)])
def test_front_scheme_reaches_the_oracle_key_field(orc, case, dfunc, nneigh):
    rng = np.random.RandomState(100 + case)
    for shape, fill, nseeds in (((9, 8, 12), 0.55, 4), ((6, 14, 7), 0.8, 3), ((10, 10, 10), 1.0, 2)):
        grid = random_blob_grid(shape, 21 + case, fill=fill, smooth=1) if fill < 1.0 else np.ones(shape, np.uint16)
        seeds = pick_seeds(grid, nseeds, 3 + case)
        if case % 2 == 1:  # two seeds on one cell (the later one wins) and a seed on an EMPTY cell
            extra = seeds[:1].copy()
            extra[0, 3] = 40
            empty = np.argwhere(grid == 0)
            rows = [seeds, extra]
            if len(empty):
                rows.append(np.uint32([[*empty[0], 41]]))
            seeds = np.concatenate(rows).astype(np.uint32)
        want = orc.flood_keys(grid.copy(), seeds, dfunc)
        got, handed_over = front_model(grid, seeds, nneigh, rng, **CASES[case])
        assert np.array_equal(got.astype(np.uint32), want), (case, shape)
        if "lose_prob" in CASES[case] or "limit" in CASES[case]:
            assert handed_over
