"""F1 parity: CUDA nearest-seed Voronoi vs the oracle (NaiveFracturer::buildCPU restated), bit-exact labels."""
import numpy as np
import pytest

from conftest import pick_seeds, random_blob_grid

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import voxelfragmentml_b200 as vf

    c = vf.Context(0)
    yield c
    c.close()


def _run_naive(ctx, grid, seeds, dfunc):
    import voxelfragmentml_b200 as vf

    g = vf.RegularGrid(ctx, grid.shape)
    g.updateSSBO(grid)
    f = vf.NaiveFracturer()
    assert f.setDistanceFunction(dfunc)
    f.build(g, seeds)
    out = g.updateGrid()
    g.close()
    return out


@pytest.mark.parametrize("dfunc", [0, 1, 2])
def test_vessel_fixture_8_seeds(ctx, orc, vessel_grid, dfunc):
    """cfg1-like: the reference's real 128x110x128 vessel, 8 OUTER seeds from the reference RNG stream (seed 80)."""
    seeds, _ = orc.seed_uniform(orc.Rng(80), vessel_grid, 8)
    want = orc.naive(vessel_grid.copy(), seeds, dfunc)
    got = _run_naive(ctx, vessel_grid, seeds, dfunc)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("shape", [(40, 36, 64), (33, 21, 76), (17, 9, 20), (13, 11, 7), (5, 6, 130), (64, 64, 8)])
@pytest.mark.parametrize("dfunc", [0, 1, 2])
def test_ragged_shapes(ctx, orc, shape, dfunc):
    """Z % 8 == 0 (128-bit path), Z % 8 == 4 (64-bit path), odd Z (generic path), partial bricks on every axis."""
    g = random_blob_grid(shape, 3)
    g[0, 0, 0] = 1
    g[-1, -1, -1] = 1
    seeds = pick_seeds(g, 7, 5)
    want = orc.naive(g.copy(), seeds, dfunc)
    got = _run_naive(ctx, g, seeds, dfunc)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("nseeds", [1, 33, 100, 300])
def test_seed_counts_and_slot_overflow(ctx, orc, nseeds):
    """Many seeds on a small grid push bricks past the 64-slot candidate list into the exact scan path; ties everywhere."""
    g = np.ones((24, 24, 64), np.uint16)
    seeds = pick_seeds(g, nseeds, nseeds)
    for dfunc in (0, 1, 2):
        want = orc.naive(g.copy(), seeds, dfunc)
        got = _run_naive(ctx, g, seeds, dfunc)
        assert np.array_equal(got, want)


def test_ties_go_to_lowest_seed_index(ctx, orc):
    g = np.ones((9, 9, 16), np.uint16)
    seeds = np.array([[4, 4, 3, 7], [4, 4, 11, 5], [4, 4, 7, 9]], np.uint32)  # labels deliberately not in index order
    for dfunc in (0, 1, 2):
        want = orc.naive(g.copy(), seeds, dfunc)
        got = _run_naive(ctx, g, seeds, dfunc)
        assert np.array_equal(got, want)
        assert got[4, 4, 5] == 7 and got[4, 4, 9] == 5  # equidistant cells: index 0 beats index 2, index 1 beats index 2


def test_prelabelled_and_empty_cells(ctx, orc):
    g = random_blob_grid((20, 20, 24), 9) * 5  # non-EMPTY cells carry an arbitrary previous label
    seeds = pick_seeds(g, 4, 1)
    got = _run_naive(ctx, g, seeds, 0)
    assert np.array_equal(got, orc.naive(g.copy(), seeds, 0))
    assert np.array_equal(got == 0, g == 0)
    empty = np.zeros((8, 8, 8), np.uint16)
    assert not _run_naive(ctx, empty, np.array([[1, 1, 1, 2]], np.uint32), 0).any()


def test_seed_outside_grid_is_rejected(ctx):
    import voxelfragmentml_b200 as vf

    g = vf.RegularGrid(ctx, (8, 8, 8))
    with pytest.raises(vf.VoxFragError) as e:
        vf.NaiveFracturer().build(g, np.array([[8, 0, 0, 2]], np.uint32))
    assert e.value.status == 1
    f = vf.NaiveFracturer()
    assert not f.setDistanceFunction(7)


def test_full_size_512_dense_sampled_and_idempotent(ctx):
    """cfg3 size: 512^3 dense, 64 seeds, Euclidean.  Checked on 200k sampled voxels against float32 brute force, plus
    idempotence (labelling a labelled grid changes nothing) and the culling-free property label in seed labels."""
    import voxelfragmentml_b200 as vf

    n = 512
    rs = np.random.RandomState(80)
    pts = rs.randint(0, n - 1, size=(64, 3))
    pts = pts[np.lexsort((pts[:, 2], pts[:, 1], pts[:, 0]))]
    seeds = np.concatenate([pts, np.arange(2, 66)[:, None]], 1).astype(np.uint32)
    g = vf.RegularGrid(ctx, (n, n, n))
    g.fillValue(1)
    f = vf.NaiveFracturer()
    f.build(g, seeds)
    a = g.updateGrid()
    f.build(g, seeds)
    b = g.updateGrid()
    assert np.array_equal(a, b)
    q = rs.randint(0, n, size=(200000, 3))
    d = q[:, None, :].astype(np.float32) - seeds[None, :, :3].astype(np.float32)
    dist = np.sqrt((d * d).sum(-1, dtype=np.float32), dtype=np.float32)
    want = seeds[np.argmin(dist, axis=1), 3]  # argmin returns the first minimum == lowest seed index
    assert np.array_equal(a[q[:, 0], q[:, 1], q[:, 2]], want.astype(np.uint16))
    assert a.min() >= 2 and a.max() <= 65
    g.close()
