"""f2 parity: CUDA marching cubes + vertex fusion + Laplacian smoothing (vf_marching_cubes) vs the oracle — identical vertex bits,
identical faces, under the deterministic ordering both define (the reference's own order is an atomicAdd race)."""
import numpy as np
import pytest

from conftest import random_blob_grid

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import voxelfragmentml_b200 as vf

    c = vf.Context(0)
    yield c
    c.close()


def _mesh(ctx, lab, target, mn, mx, **kw):
    import voxelfragmentml_b200 as vf

    g = vf.RegularGrid(ctx, lab.shape)
    g.setAABB(mn, mx, lab.shape)
    g.updateSSBO(lab)
    out = g.triangulateField(target, **kw)
    g.close()
    return out


def _same(a, b):
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_fragments_of_the_labelled_vessel(ctx, orc, vessel_grid):
    seeds, _ = orc.seed_uniform(orc.Rng(80), vessel_grid, 6)
    lab = orc.detect_boundaries(orc.naive(vessel_grid.copy(), seeds, 0), 1)
    mn, mx = np.float32([-0.5, -0.43, -0.5]), np.float32([0.5, 0.43, 0.5])
    for target in (2, 4, 7):
        wv, wf = orc.marching_cubes(lab, target, mn, mx)
        gv, gf = _mesh(ctx, lab, target, mn, mx)
        assert len(wf) > 1000 and np.array_equal(gf, wf)
        assert _same(gv, wv)


@pytest.mark.parametrize("shape", [(28, 24, 30), (9, 17, 5), (40, 6, 33), (3, 3, 3)])
def test_random_blobs_with_and_without_smoothing(ctx, orc, shape):
    occ = (random_blob_grid(shape, 5) != 0).astype(np.uint16)
    rs = np.random.RandomState(2)
    lab = (occ * rs.randint(2, 5, size=shape)).astype(np.uint16)   # noisy labels: every case of the table shows up
    lab = orc.detect_boundaries(lab, 1)
    mn, mx = np.float32([-0.4, -0.3, -0.5]), np.float32([0.4, 0.3, 0.5])
    for target in (2, 3):
        for iters in (0.0, 0.3):
            wv, wf = orc.marching_cubes(lab, target, mn, mx, nb_iters=int(np.float32(max(shape)) * np.float32(iters)), b_iters=int(np.float32(max(shape)) * np.float32(iters)))
            gv, gf = _mesh(ctx, lab, target, mn, mx, boundaryMCIterations=iters, nonBoundaryMCIterations=iters)
            assert np.array_equal(gf, wf) and _same(gv, wv), (shape, target, iters)


def test_absent_label_full_grid_and_bad_arguments(ctx, orc):
    import voxelfragmentml_b200 as vf

    g = np.full((6, 4, 5), 2, np.uint16)
    v, f = _mesh(ctx, g, 9, np.float32([0, 0, 0]), np.float32([6, 4, 5]))
    assert len(v) == 0 and len(f) == 0
    wv, wf = orc.marching_cubes(g, 2, np.float32([0, 0, 0]), np.float32([6, 4, 5]))
    gv, gf = _mesh(ctx, g, 2, np.float32([0, 0, 0]), np.float32([6, 4, 5]))
    assert np.array_equal(gf, wf) and _same(gv, wv)
    grid = vf.RegularGrid(ctx, (4, 4, 4))
    with pytest.raises(vf.VoxFragError):
        grid.triangulateField(1)      # FREE is not a fragment label
    grid.close()
    # _marchingCubesSubdivisions is a dead parameter of the reference (FractureParameters.h:56; RegularGrid.cpp:423 passes a literal 1): same mesh whatever it says
    sv, sf = _mesh(ctx, g, 2, np.float32([0, 0, 0]), np.float32([6, 4, 5]), marchingCubesSubdivisions=3)
    assert np.array_equal(sf, wf) and _same(sv, wv)
