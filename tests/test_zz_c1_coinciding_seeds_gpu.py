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
    assert np.array_equal(_c1(ctx, lab, seeds), want)


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
