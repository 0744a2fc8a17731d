"""What csrc/ccl.cu computes for C1, restated in numpy (scipy.ndimage components) and checked against the oracle — i.e. against the reference's
removeIsolatedRegionsCPU compiled in place — on seed lists where several seeds share a cell:
    plant every seed's label on its cell, the later seed winning a shared cell (NaiveFracturer.cpp:120-123);
    keep the 6-connected whole-word components that hold a START: a seed's cell when it carries the seed's label, otherwise — a later seed took
    the cell — the neighbours of the cell that carry the seed's label (the reference searches from every seed with the seed's own label, :127-146).
The second clause was missing before the end of round 1: with CADScene's seed copies (CADScene.cpp:651) whole regions lost their start."""
import numpy as np
import pytest
from scipy import ndimage

from conftest import pick_seeds, random_blob_grid

NB = [(1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1)]


def c1_model(lab, seeds, displaced_seeds_start_from_neighbours=True):
    g = lab.copy()
    for x, y, z, w in seeds:
        g[x, y, z] = w
    comp = np.zeros(g.shape, np.int64)
    base = 0
    for L in np.unique(g[g > 1]):
        c, n = ndimage.label(g == L)
        comp[c > 0] = c[c > 0] + base
        base += n
    keep = set()
    for x, y, z, w in (tuple(int(v) for v in s) for s in seeds):
        if g[x, y, z] == w or not displaced_seeds_start_from_neighbours:
            keep.add(int(comp[x, y, z]))
        else:
            for d in NB:
                n = (x + d[0], y + d[1], z + d[2])
                if all(0 <= n[k] < g.shape[k] for k in range(3)) and g[n] == w:
                    keep.add(int(comp[n]))
    keep.discard(0)
    return np.where(np.isin(comp, list(keep)), g, 0).astype(np.uint16)


@pytest.mark.parametrize("nf,ne,dfunc", [(6, 12, 0), (4, 8, 1)])
def test_seed_copies_of_fracture_model_keep_their_regions(orc, vessel_grid, nf, ne, dfunc):
    seeds = orc.make_seeds(orc.Rng(80 + nf), vessel_grid, nf, ne, merge_dfunc=0)
    lab = orc.naive(vessel_grid.copy(), seeds, dfunc)
    want = orc.remove_isolated_regions_cpu(lab.copy(), seeds)
    assert np.array_equal(c1_model(lab, seeds), want)
    # the rule before the fix: the displaced originals lose their regions
    assert (c1_model(lab, seeds, displaced_seeds_start_from_neighbours=False) != want).sum() > 10000


@pytest.mark.parametrize("k", range(3))
def test_copies_and_extras_on_porous_blobs(orc, k):
    b = random_blob_grid((30, 28, 40), k, fill=0.55, smooth=1)
    sd = pick_seeds(b, 6, k)
    copies = sd.copy()
    copies[:, 3] = sd[:, 3] | 0x100
    more = pick_seeds(b, 5, 50 + k)
    more[:, 3] = 0x200 | (2 + np.arange(5) % 6)
    full = np.concatenate([sd, copies, more]).astype(np.uint32)
    lab = orc.naive(b.copy(), full, k % 3)
    assert np.array_equal(c1_model(lab, full), orc.remove_isolated_regions_cpu(lab.copy(), full))
