"""V2 parity: CUDA binned SAT voxelizer vs the oracle's restatement of Intersections3D::intersect — bit-exact occupancy
(the 1e-6-relative margin report of BASELINE.json is therefore a report of zero mismatches; it is still computed)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import voxelfragmentml_b200 as vf

    c = vf.Context(0)
    yield c
    c.close()


def _voxelize(ctx, verts, faces, mn, mx, dims):
    import voxelfragmentml_b200 as vf

    g = vf.RegularGrid(ctx, dims)
    g.setAABB(mn, mx, dims)
    g.fill(verts, faces)
    out = g.updateGrid()
    g.close()
    return out


@pytest.mark.parametrize("maxvox", [64, 128])
def test_vessel_mesh_cfg1(ctx, orc, maxvox):
    from voxelfragmentml_b200 import synth

    v, f = synth.vessel_mesh(0)
    assert 19000 < len(f) < 21000
    mn, mx = synth.mesh_aabb(v)
    dims = orc.dims_rule(mn, mx, maxvox)
    want, margin = orc.voxelize_sat(v, f, mn, mx, dims, want_margin=True)
    got = _voxelize(ctx, v, f, mn, mx, dims)
    mismatch = got != want
    near = margin < 1e-6
    assert not (mismatch & ~near).any(), "occupancy differs outside the 1e-6-relative SAT margin"
    assert not mismatch.any(), f"mismatch fraction {mismatch.mean():.3e} (expected exactly 0: same float32 op order, no FMA)"
    assert 0.005 < (got != 0).mean() < 0.2 and set(np.unique(got)) == {0, 1}


def test_random_triangle_soup_ragged_dims(ctx, orc):
    rs = np.random.RandomState(4)
    for dims in [(33, 21, 70), (8, 4, 32), (5, 7, 9), (40, 40, 40)]:
        v = rs.uniform(-0.5, 0.5, size=(90, 3)).astype(np.float32)
        f = rs.randint(0, 90, size=(150, 3)).astype(np.uint32)   # includes degenerate and huge triangles
        mn, mx = np.float32([-0.5, -0.45, -0.5]), np.float32([0.5, 0.45, 0.48])
        want = orc.voxelize_sat(v, f, mn, mx, dims)
        got = _voxelize(ctx, v, f, mn, mx, dims)
        assert np.array_equal(got, want), dims


def test_axis_aligned_faces_on_cell_boundaries(ctx, orc):
    """Planes exactly on voxel faces exercise every >= / > tie of the predicate."""
    v = np.float32([[-0.5, -0.5, 0.0], [0.5, -0.5, 0.0], [0.5, 0.5, 0.0], [-0.5, 0.5, 0.0],
                    [0.25, -0.5, -0.5], [0.25, 0.5, -0.5], [0.25, 0.5, 0.5], [0.25, -0.5, 0.5]])
    f = np.uint32([[0, 1, 2], [0, 2, 3], [4, 5, 6], [4, 6, 7]])
    mn, mx = np.float32([-0.5] * 3), np.float32([0.5] * 3)
    for dims in [(16, 16, 16), (8, 12, 20)]:
        want = orc.voxelize_sat(v, f, mn, mx, dims)
        got = _voxelize(ctx, v, f, mn, mx, dims)
        assert np.array_equal(got, want)


def test_invalid_face_index_and_refill_clears(ctx, orc):
    import voxelfragmentml_b200 as vf

    g = vf.RegularGrid(ctx, (16, 16, 16))
    g.fillValue(7)
    v = np.float32([[0, 0, 0], [0.2, 0, 0], [0, 0.2, 0]])
    with pytest.raises(vf.VoxFragError):
        g.fill(v, np.uint32([[0, 1, 3]]))
    g.fill(v, np.uint32([[0, 1, 2]]))
    out = g.updateGrid()
    assert set(np.unique(out)) == {0, 1}
    g.close()


def test_cfg1_pipeline_end_to_end(ctx, orc):
    """BASELINE cfg1: vessel -> SAT voxelize at 128 -> 8 OUTER seeds (seed 80) -> NAIVE EUCLIDEAN -> detectBoundaries -> undoMask."""
    import voxelfragmentml_b200 as vf
    from voxelfragmentml_b200 import synth

    v, f = synth.vessel_mesh(0)
    mn, mx = synth.mesh_aabb(v)
    dims = orc.dims_rule(mn, mx, 128)
    lib_dims = np.zeros(3, np.uint32)
    vf._capi.load().vf_dims_rule(mn.ctypes.data, mx.ctypes.data, 128, lib_dims.ctypes.data)
    assert tuple(lib_dims) == dims
    g = vf.RegularGrid(ctx, dims)
    g.setAABB(mn, mx, dims)
    g.fill(v, f)
    ctx.initSeed(80)
    p = vf.FractureParameters(_fractureAlgorithm=vf.FractureAlgorithm.NAIVE, _distanceFunction=vf.DistanceFunction.EUCLIDEAN, _numSeeds=8,
                              _numExtraSeeds=0, _removeIsolatedRegions=0)
    seeds, _ = vf.fracture_model(g, p)
    g.undoMask()
    got = g.updateGrid()
    occ = orc.voxelize_sat(v, f, mn, mx, dims)
    wseeds = orc.make_seeds(orc.Rng(80), occ, 8, 0)
    want = orc.undo_mask(orc.detect_boundaries(orc.naive(occ.copy(), wseeds, 0), 1), 15, False)
    assert np.array_equal(seeds, wseeds) and np.array_equal(got, want)
    g.close()
