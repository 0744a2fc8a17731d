"""V1 parity: CUDA Tetravoxelizer occupancy (vf_voxelize_solid: XOR-ed tetrahedron slices, RegularGrid::fill as the reference runs it
today) vs the oracle — bit-exact (same float32 shader arithmetic, same fixed-point coverage rule).  The reference's own coverage
is decided by the GL rasteriser, so this row is parity-unpinned by construction; tests/test_solid_cpu.py checks the oracle
against an independent ray-parity computation."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import voxelfragmentml_b200 as vf

    c = vf.Context(0)
    yield c
    c.close()


def _solid(ctx, verts, faces, mn, mx, dims):
    import voxelfragmentml_b200 as vf

    g = vf.RegularGrid(ctx, dims)
    g.setAABB(mn, mx, dims)
    g.fillValue(9)  # fill() starts from a cleaned grid
    occ = g.fillSolid(verts, faces)
    out = g.updateGrid()
    g.close()
    return out, occ


@pytest.mark.parametrize("maxvox", [64, 128, 256])
def test_vessel_mesh_solid(ctx, orc, maxvox):
    from voxelfragmentml_b200 import synth

    v, f = synth.vessel_mesh(maxvox % 3)
    mn, mx = synth.mesh_aabb(v)
    dims = orc.dims_rule(mn, mx, maxvox)
    want = orc.voxelize_solid(v, f, mn, mx, dims)
    got, occ = _solid(ctx, v, f, mn, mx, dims)
    assert np.array_equal(got, want)
    assert occ == int(want.sum()) and 0.05 < want.mean() < 0.5
    assert not got[:, 0, :].any()  # slice 0 samples the AABB's bottom plane: nothing is strictly below it


def test_random_closed_and_open_soups_ragged_dims(ctx, orc):
    """Random triangles (open surfaces: the XOR parity is still well defined), huge and degenerate ones, dims with Z % 8 != 0,
    a mesh that sticks out of the grid AABB, and bands (X large enough to be cut into several CTAs per slice)."""
    rs = np.random.RandomState(12)
    for dims in [(33, 21, 70), (8, 4, 32), (5, 7, 9), (40, 40, 40), (300, 6, 12), (16, 300, 20)]:
        v = rs.uniform(-0.6, 0.6, size=(60, 3)).astype(np.float32)
        f = rs.randint(0, 60, size=(110, 3)).astype(np.uint32)
        mn, mx = np.float32([-0.5, -0.45, -0.5]), np.float32([0.5, 0.45, 0.48])
        want = orc.voxelize_solid(v, f, mn, mx, dims)
        got, occ = _solid(ctx, v, f, mn, mx, dims)
        assert np.array_equal(got, want), dims
        assert occ == int(want.sum()) and occ > 0


def test_cube_faces_on_cell_boundaries(ctx, orc):
    """An axis-aligned box whose faces lie exactly on cell boundaries and pixel centres: every tie of the coverage rule."""
    c = np.float32([[x, y, z] for x in (-0.25, 0.25) for y in (-0.25, 0.25) for z in (-0.25, 0.25)])
    quads = [(0, 1, 3, 2), (4, 6, 7, 5), (0, 4, 5, 1), (2, 3, 7, 6), (0, 2, 6, 4), (1, 5, 7, 3)]
    f = np.uint32([t for q in quads for t in ((q[0], q[1], q[2]), (q[0], q[2], q[3]))])
    mn, mx = np.float32([-0.5] * 3), np.float32([0.5] * 3)
    for dims in [(16, 16, 16), (8, 12, 20), (17, 9, 31)]:
        want = orc.voxelize_solid(c, f, mn, mx, dims)
        got, _ = _solid(ctx, c, f, mn, mx, dims)
        assert np.array_equal(got, want), dims
        assert want.any()


def test_solid_invalid_face_index_and_empty_mesh(ctx):
    import voxelfragmentml_b200 as vf

    g = vf.RegularGrid(ctx, (16, 16, 16))
    v = np.float32([[0, 0, 0], [0.2, 0, 0], [0, 0.2, 0]])
    with pytest.raises(vf.VoxFragError):
        g.fillSolid(v, np.uint32([[0, 1, 3]]))
    g.fillValue(5)
    assert g.fillSolid(v, np.zeros((0, 3), np.uint32)) == 0
    assert not g.updateGrid().any()
    g.close()
