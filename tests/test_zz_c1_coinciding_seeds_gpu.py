"""C1 when several seeds share a cell — the seed list CADScene::fractureModel builds with extra seeds holds a copy of every original seed on the
original's cell (CADScene.cpp:647-655), and NaiveFracturer::build hands that list to removeIsolatedRegions.  The reference searches from
EVERY seed with the seed's own label (NaiveFracturer.cpp:116-146), so a seed whose cell a later seed took still keeps its region.  Found late in
round 1 (the union-find path marked only the component of the cell's final label); the marking rule was extended in csrc/ccl.cu and is
checked here.  The file sorts last on purpose: it is the newest, least exercised check."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _c1(ctx, lab, seeds):
    import voxelfragmentml_b200 as vf

    g = vf.RegularGrid(ctx, lab.shape)
    g.updateSSBO(lab)
    vf.NaiveFracturer.removeIsolatedRegions(g, seeds)
    out = g.updateGrid()
    g.close()
    return out


@pytest.fixture(scope="module")
def ctx():
    import voxelfragmentml_b200 as vf

    c = vf.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("nf,ne,dfunc", [(6, 12, 0), (4, 8, 1), (8, 16, 2)])
def test_c1_with_the_seed_copies_of_fracture_model(ctx, orc, vessel_grid, nf, ne, dfunc):
    seeds = orc.make_seeds(orc.Rng(80 + nf), vessel_grid, nf, ne, merge_dfunc=0)  # [originals, copies on the same cells, extras]
    lab = orc.naive(vessel_grid.copy(), seeds, dfunc)
    want = orc.remove_isolated_regions_cpu(lab.copy(), seeds)
    assert (want > 1).sum() > 100000  # the regions survive: every original seed's cell now carries its copy's label
    for mode in (2, 0):  # the descent certificate (2: on a grid of any size), then the default policy (the union-find on a grid this small)
        ctx.setC1Mode(mode)
        try:
            assert np.array_equal(_c1(ctx, lab, seeds), want), f"c1 mode {mode}"
        finally:
            ctx.setC1Mode(0)


def test_fracture_model_naive_with_extra_seeds(ctx, orc, vessel_grid):
    import voxelfragmentml_b200 as vf

    p = vf.FractureParameters(_fractureAlgorithm=vf.FractureAlgorithm.NAIVE, _distanceFunction=vf.DistanceFunction.EUCLIDEAN, _numSeeds=6,
                              _numExtraSeeds=12, _erode=0)
    g = vf.RegularGrid(ctx, vessel_grid.shape)
    g.updateSSBO(vessel_grid)
    ctx.initSeed(80)
    seeds, _ = vf.fracture_model(g, p)
    wseeds = orc.make_seeds(orc.Rng(80), vessel_grid, 6, 12, merge_dfunc=0)
    want = orc.remove_isolated_regions_cpu(orc.naive(vessel_grid.copy(), wseeds, 0), wseeds)
    want = orc.detect_boundaries(want, 1)
    assert np.array_equal(seeds, wseeds) and np.array_equal(g.updateGrid(), want)
    g.close()


def _adversarial_seed_lists(orc, lab, seeds, rs):
    """seed lists the reference code pins through the oracle (tests/test_oracle_vs_ref.py): several seeds per cell, seeds on EMPTY and on
    foreign-label cells, displaced seeds whose neighbours lie in other tiles / segments"""
    out = []
    s3 = np.concatenate([seeds, seeds[:4], seeds[:4]]).astype(np.uint32)  # three seeds on one cell, three labels
    s3[len(seeds):len(seeds) + 4, 3] = seeds[:4, 3] | 0x100
    s3[len(seeds) + 4:, 3] = seeds[:4, 3] | 0x200
    out.append(s3)
    empty = np.argwhere(lab == 0)
    if len(empty):
        e = seeds.copy()
        e[0, :3] = empty[rs.randint(len(empty))]  # a seed on an EMPTY cell: its label is planted there
        out.append(e)
    foreign = seeds.copy()
    other = np.argwhere(lab == seeds[1, 3])
    foreign[0, :3] = other[rs.randint(len(other))]  # a seed on a cell another seed labelled
    out.append(foreign)
    # copies placed on cells at tile / segment borders (x, y multiples of 16, z multiples of 32 and their predecessors)
    edge = np.argwhere((lab > 1) & ((np.indices(lab.shape)[0] % 16 >= 15) | (np.indices(lab.shape)[1] % 16 == 0) | (np.indices(lab.shape)[2] % 32 >= 31)))
    if len(edge) >= 6:
        pick = edge[rs.choice(len(edge), 6, replace=False)]
        a = np.concatenate([pick, lab[tuple(pick.T)][:, None]], 1)
        b = a.copy()
        b[:, 3] = 0x300 | (2 + np.arange(6))
        out.append(np.concatenate([seeds, a, b]).astype(np.uint32))  # the later copies displace seeds that own a region across the border
    return out


@pytest.mark.parametrize("c1_mode", [2, 1, 0])  # 2: the descent certificate on grids of any size, 1: the union-find, 0: the default policy
def test_c1_adversarial_seed_lists(ctx, orc, vessel_grid, c1_mode):
    from conftest import pick_seeds, random_blob_grid

    rs = np.random.RandomState(7)
    ctx.setC1Mode(c1_mode)
    try:
        cases = [(vessel_grid, pick_seeds(vessel_grid, 9, 5), 0)]
        cases += [(random_blob_grid((45, 37, 72), 20 + k, fill=0.5 + 0.05 * k, smooth=1), None, k % 3) for k in range(3)]
        dense = np.ones((48, 40, 64), np.uint16)
        cases.append((dense, pick_seeds(dense, 12, 2), 1))
        for grid, seeds, dfunc in cases:
            if seeds is None:
                seeds = pick_seeds(grid, 7, 3)
            lab = orc.naive(grid.copy(), seeds, dfunc)
            for sl in _adversarial_seed_lists(orc, lab, seeds, rs):
                want = orc.remove_isolated_regions_cpu(lab.copy(), sl)
                got = _c1(ctx, lab, sl)
                assert np.array_equal(got, want), f"mode {c1_mode}: {int((got != want).sum())} cells differ"
    finally:
        ctx.setC1Mode(0)
