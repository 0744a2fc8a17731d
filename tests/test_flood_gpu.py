"""F2/F3/C1 parity: CUDA tile-worklist flood and connected-to-seed cleanup vs the oracle, bit-exact labels."""
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


def _run_flood(ctx, grid, seeds, dfunc, id_bits=0):
    import voxelfragmentml_b200 as vf

    g = vf.RegularGrid(ctx, grid.shape)
    g.updateSSBO(grid)
    f = vf.FloodFracturer()
    assert f.setDistanceFunction(dfunc)
    f.build(g, seeds, id_bits=id_bits)
    out = g.updateGrid()
    g.close()
    return out, f.last_stats


@pytest.mark.parametrize("dfunc", [1, 2, 0])
def test_vessel_fixture_16_seeds(ctx, orc, vessel_grid, dfunc):
    seeds, _ = orc.seed_uniform(orc.Rng(80), vessel_grid, 16)
    want, st = orc.flood(vessel_grid.copy(), seeds, dfunc)
    got, gst = _run_flood(ctx, vessel_grid, seeds, dfunc)
    assert np.array_equal(got, want)
    assert gst.max_dist == st.max_dist and gst.disjoint_rounds == 1 and gst.freed_voxels == 0
    assert gst.front_levels >= 1 or (gst.tile_rounds >= 1 and gst.tile_visits >= gst.tile_rounds)


@pytest.mark.parametrize("shape", [(40, 36, 64), (33, 21, 76), (17, 19, 35), (13, 11, 7), (5, 6, 130), (70, 9, 33)])
@pytest.mark.parametrize("dfunc", [1, 2])
def test_ragged_shapes_and_partial_tiles(ctx, orc, shape, dfunc):
    g = random_blob_grid(shape, 11, fill=0.5, smooth=1)
    seeds = pick_seeds(g, 6, 2)
    want, _ = orc.flood(g.copy(), seeds, dfunc)
    got, _ = _run_flood(ctx, g, seeds, dfunc)
    assert np.array_equal(got, want)


def test_labyrinth_long_geodesics(ctx, orc):
    """A serpentine corridor: geodesic distance ~ 20x the grid diameter, crosses every tile many times."""
    g = np.zeros((48, 40, 96), np.uint16)
    for x in range(0, 48, 2):
        g[x, :, :] = 1
    for i, x in enumerate(range(1, 47, 2)):
        g[x, 39 if i % 2 == 0 else 0, :] = 1
    seeds = np.array([[0, 0, 0, 2], [46, 0, 95, 3], [24, 20, 50, 4]], np.uint32)
    for dfunc in (1, 2):
        want, st = orc.flood(g.copy(), seeds, dfunc)
        got, gst = _run_flood(ctx, g, seeds, dfunc)
        assert np.array_equal(got, want)
        assert gst.max_dist == st.max_dist and st.max_dist > 200


def test_unreachable_cells_stay_free_and_seed_on_empty_cell(ctx, orc):
    g = np.zeros((20, 20, 40), np.uint16)
    g[2:8, 2:8, 2:30] = 1
    g[12:18, 12:18, 5:35] = 1      # island without a seed: must stay FREE (1)
    seeds = np.array([[3, 3, 3, 2], [7, 7, 29, 3], [10, 10, 10, 4]], np.uint32)  # third seed sits on an EMPTY cell
    want, _ = orc.flood(g.copy(), seeds, 1)
    got, _ = _run_flood(ctx, g, seeds, 1)
    assert np.array_equal(got, want)
    assert (got[12:18, 12:18, 5:35] == 1).all() and got[10, 10, 10] == 4


@pytest.mark.parametrize("dfunc", [1, 2])
def test_extra_seeds_disjoint_rounds(ctx, orc, vessel_grid, dfunc):
    """F3: numExtraSeeds = 2*numSeeds as in dataset generation (FragmentationProcedure.h:12-13, CADScene.cpp:298-299)."""
    for trial, nf in enumerate([2, 4, 7, 10]):
        seeds = orc.make_seeds(orc.Rng(80 + trial), vessel_grid, nf, 2 * nf, merge_dfunc=0)
        want, st = orc.flood(vessel_grid.copy(), seeds, dfunc)
        got, gst = _run_flood(ctx, vessel_grid, seeds, dfunc)
        assert np.array_equal(got, want), f"nf={nf}"
        assert gst.freed_voxels == st.freed_voxels and gst.disjoint_rounds == st.rounds
        assert set(np.unique(got)) <= set(range(0, nf + 2))


@pytest.mark.parametrize("dfunc", [1, 2])
def test_orphaned_extra_seed_is_dissolved_and_reflooded(ctx, orc, dfunc):
    """U-shaped corridor: an extra seed that is Euclidean-nearest to fragment 2 but geodesically behind fragment 3 grows a
    region that is not connected to fragment 2's own; the disjoint step frees it and the re-flood hands it to fragment 3."""
    g = np.zeros((5, 9, 100), np.uint16)
    g[1:4, 1:4, :] = 1          # outbound corridor
    g[1:4, 5:8, :] = 1          # return corridor
    g[1:4, 1:8, 96:100] = 1     # the turn
    frags = np.array([[2, 2, 3, 2], [2, 4, 98, 3]], np.uint32)
    extra = np.array([[2, 6, 3, 0]], np.uint32)
    seeds = np.concatenate([frags, orc.merge_seeds(frags, np.concatenate([frags, extra]), 0)])
    assert list(seeds[:, 3]) == [2, 3, 2 | 1 << 8, 3 | 1 << 8, 2 | 2 << 8]
    want, st = orc.flood(g.copy(), seeds, dfunc)
    got, gst = _run_flood(ctx, g, seeds, dfunc)
    assert st.rounds == 2 and st.freed_voxels > 100
    assert np.array_equal(got, want)
    assert gst.disjoint_rounds == 2 and gst.freed_voxels == st.freed_voxels
    assert got[2, 6, 3] == 3 and got[2, 2, 3] == 2


def test_extra_seeds_on_random_blobs(ctx, orc):
    for trial in range(6):
        g = random_blob_grid((37, 29, 45), 200 + trial, fill=0.5, smooth=1)
        seeds = orc.make_seeds(orc.Rng(trial), g, 3, 9, merge_dfunc=trial % 3)
        for dfunc in (1, 2):
            want, st = orc.flood(g.copy(), seeds, dfunc)
            got, gst = _run_flood(ctx, g, seeds, dfunc)
            assert np.array_equal(got, want)
            assert gst.freed_voxels == st.freed_voxels


@pytest.mark.parametrize("fill", [0.35, 0.5, 0.7])
def test_extra_seeds_on_unsmoothed_noise_across_tiles(ctx, orc, fill):
    """Raw occupancy noise on a grid of several tiles per axis: most contacts between fragments' cells are diagonal-only and runs are
    one or two cells long, which is the worst case for the component step of F3 (diagonal-pair filtering, witness skipping across tile
    borders under the 26-neighbourhood) and for the flood's dealt-candidate steps."""
    for trial in range(3):
        g = random_blob_grid((40, 36, 72), 900 + trial, fill=fill, smooth=0)
        seeds = orc.make_seeds(orc.Rng(10 + trial), g, 4, 20, merge_dfunc=trial % 3)
        for dfunc in (2, 1):
            want, st = orc.flood(g.copy(), seeds, dfunc)
            got, gst = _run_flood(ctx, g, seeds, dfunc)
            assert np.array_equal(got, want), (fill, trial, dfunc)
            assert gst.freed_voxels == st.freed_voxels


def test_id_bits_15_many_labels(ctx, orc):
    """cfg5's label space: 256+ seeds need ids beyond 8 bits (SURVEY finding 7) -> whole word is the id, no prefixes."""
    g = random_blob_grid((64, 64, 64), 5, fill=0.6, smooth=1)
    seeds = pick_seeds(g, 300, 9)
    want, _ = orc.flood(g.copy(), seeds, 1, id_bits=15)
    got, _ = _run_flood(ctx, g, seeds, 1, id_bits=15)
    assert np.array_equal(got, want) and got.max() == 301


def test_full_size_256_upsampled_vessel(ctx, orc, vessel_grid):
    """cfg2 size: the reference vessel upsampled x2 (256x220x256, 889k*8 occupied cells), Manhattan, 16 seeds."""
    big = np.ascontiguousarray(np.kron(vessel_grid, np.ones((2, 2, 2), np.uint16)))
    seeds, _ = orc.seed_uniform(orc.Rng(80), big, 16)
    want, st = orc.flood(big.copy(), seeds, 1)
    got, gst = _run_flood(ctx, big, seeds, 1)
    assert np.array_equal(got, want) and gst.max_dist == st.max_dist
    # running twice on the labelled grid re-homogenises and gives the same result (FloodFracturer.cpp:99)
    import voxelfragmentml_b200 as vf

    g = vf.RegularGrid(ctx, big.shape)
    g.updateSSBO(got)
    f = vf.FloodFracturer()
    f.build(g, seeds)
    assert np.array_equal(g.updateGrid(), want)
    g.close()


# ------------------------------------------------------------------------------------------------ C1
def _run_c1(ctx, grid, seeds):
    import voxelfragmentml_b200 as vf

    g = vf.RegularGrid(ctx, grid.shape)
    g.updateSSBO(grid)
    vf.NaiveFracturer.removeIsolatedRegions(g, seeds)
    out = g.updateGrid()
    g.close()
    return out


@pytest.mark.parametrize("dfunc", [0, 1, 2])
def test_c1_after_naive_on_vessel(ctx, orc, vessel_grid, dfunc):
    seeds, _ = orc.seed_uniform(orc.Rng(80), vessel_grid, 12)
    lab = orc.naive(vessel_grid.copy(), seeds, dfunc)
    want = orc.remove_isolated_regions_cpu(lab.copy(), seeds)
    got = _run_c1(ctx, lab, seeds)
    assert np.array_equal(got, want)


def test_c1_islands_free_cells_and_foreign_seed_cell(ctx, orc):
    for trial in range(5):
        g = random_blob_grid((35, 33, 70), 300 + trial, fill=0.45, smooth=1)
        seeds = pick_seeds(g, 8, trial)
        lab = orc.naive(g.copy(), seeds, 1)  # Manhattan cells are not convex: islands appear in porous blobs
        lab[1:3, 1:3, 1:3] = 1               # stray FREE cells are dropped too (output is rebuilt from an EMPTY grid)
        s2 = seeds.copy()
        s2[0, :3] = np.argwhere(lab == s2[1, 3])[0]  # seed 0 sits on a cell labelled by seed 1: its cell still becomes s2[0].w
        for sd in (seeds, s2):
            want = orc.remove_isolated_regions_cpu(lab.copy(), sd)
            got = _run_c1(ctx, lab, sd)
            assert np.array_equal(got, want)


@pytest.mark.parametrize("levels", [1, 3, 8, 40, 100000])
def test_result_does_not_depend_on_the_distance_window_or_blocking_waits(orc, vessel_grid, levels):
    """vf_ctx_set_flood_levels / vf_ctx_set_blocking_sync are scheduling knobs: labels stay bit-exact (plain and extra-seed floods)."""
    import voxelfragmentml_b200 as vf

    c = vf.Context(0)
    c.setFloodLevels(levels)
    c.setFloodFront({1: 0, 8: 0, 100000: 0, 3: 150}.get(levels, 600))  # tiles only, or taking over from the thin-front solver somewhere on the way
    c.setBlockingSync({1: 1, 3: 2}.get(levels, 0))  # spinning, sleeping on a blocking event (1), polling with sched_yield (2)
    seeds, _ = orc.seed_uniform(orc.Rng(80), vessel_grid, 12)
    for dfunc in (1, 2):
        want, _ = orc.flood(vessel_grid.copy(), seeds, dfunc)
        got, _ = _run_flood(c, vessel_grid, seeds, dfunc)
        assert np.array_equal(got, want)
    wseeds = orc.make_seeds(orc.Rng(80), vessel_grid, 6, 12, merge_dfunc=0)
    want, _ = orc.flood(vessel_grid.copy(), wseeds, orc.CHEBYSHEV)
    got, _ = _run_flood(c, vessel_grid, wseeds, 2)
    assert np.array_equal(got, want)
    c.close()


@pytest.mark.parametrize("wait", [1, 2])
def test_c1_and_histogram_under_the_other_wait_modes(orc, vessel_grid, wait):
    """vf_ctx_set_blocking_sync(1 | 2): C1 learns the certificate's verdict through a copy + blocking event, or by polling with sched_yield;
    the histogram's read-back waits the same way.  Dense grid (the certificate handles it) and the vessel shell (it declines)."""
    import voxelfragmentml_b200 as vf

    c = vf.Context(0)
    c.setBlockingSync(wait)
    c.setC1Mode(2)  # the certificate on these small grids too (the default keeps it for grids of at least 2^26 cells)
    dense = np.ones((40, 32, 48), np.uint16)
    for grid, ns in ((dense, 9), (vessel_grid, 12)):
        seeds = pick_seeds(grid, ns, 81)
        lab = orc.naive(grid.copy(), seeds, 0)
        want = orc.remove_isolated_regions_cpu(lab.copy(), seeds)
        g = vf.RegularGrid(c, lab.shape)
        g.updateSSBO(lab)
        vf.NaiveFracturer.removeIsolatedRegions(g, seeds)
        counts, occ = g.countValues()
        assert np.array_equal(g.updateGrid(), want)
        wc, wocc = orc.count_values(want)
        assert np.array_equal(np.asarray(counts), np.asarray(wc)) and int(occ) == int(wocc)
        g.close()
    c.close()


@pytest.mark.parametrize("mode", [0, 1, 2, 4])
def test_result_does_not_depend_on_how_the_round_loop_is_driven(orc, vessel_grid, mode):
    """setFloodMode: one cooperative launch per phase with 1 / 2 / 4 CTAs per SM (the round loop on the device, default 4) or one launch per
    round (0).  Plain flood, extra seeds (two phases + union-find in between), a labyrinth (hundreds of rounds), a ragged grid without TMA."""
    import voxelfragmentml_b200 as vf

    c = vf.Context(0)
    c.setFloodMode(mode)
    c.setFloodFront(0)  # the tile rounds do all the work here; test_result_does_not_depend_on_the_front_limit mixes the two solvers
    seeds, _ = orc.seed_uniform(orc.Rng(80), vessel_grid, 16)
    for dfunc in (1, 2):
        want, st = orc.flood(vessel_grid.copy(), seeds, dfunc)
        got, gst = _run_flood(c, vessel_grid, seeds, dfunc)
        assert np.array_equal(got, want) and gst.max_dist == st.max_dist
        xs = orc.make_seeds(orc.Rng(85), vessel_grid, 5, 10)
        want, st = orc.flood(vessel_grid.copy(), xs, dfunc)
        got, gst = _run_flood(c, vessel_grid, xs, dfunc)
        assert np.array_equal(got, want) and gst.disjoint_rounds == st.rounds and gst.freed_voxels == st.freed_voxels
    lab = np.zeros((48, 40, 64), np.uint16)  # serpentine corridor: the front crosses the same tiles again and again
    lab[1:-1, 1:-1, 1:-1] = 1
    for x in range(4, 44, 4):
        lab[x, (1 if (x // 4) % 2 else 4):(36 if (x // 4) % 2 else 39), :] = 0
    sd = np.uint32([[2, 2, 2, 2], [45, 37, 60, 3]])
    want, st = orc.flood(lab.copy(), sd, 1)
    got, gst = _run_flood(c, lab, sd, 1)
    assert np.array_equal(got, want) and gst.max_dist == st.max_dist
    rag = random_blob_grid((33, 21, 75), 9, fill=0.6)  # Z % 4 != 0: keys are staged row by row, the loop is driven from the host
    sd = pick_seeds(rag, 5, 2)
    want, _ = orc.flood(rag.copy(), sd, 2)
    got, _ = _run_flood(c, rag, sd, 2)
    assert np.array_equal(got, want)
    c.close()


@pytest.mark.parametrize("c1_mode", [2, 1, 0])
def test_c1_both_formulations_are_bit_exact(ctx, orc, vessel_grid, c1_mode):
    """C1 by descent certificate (csrc/c1_descent.cu: certificate pass + list work; setC1Mode(2) = on grids of any size, the default 0 keeps it for
    grids of at least 2^26 cells) and by the union-find alone (setC1Mode(1)); same cases as the C1 tests above plus a dense Voronoi grid (few
    failing cells), labels without a seed and a grid whose Z is not a multiple of 8 (the certificate declines, the union-find runs)."""
    ctx.setC1Mode(c1_mode)
    try:
        _c1_cases(ctx, orc, vessel_grid)
    finally:
        ctx.setC1Mode(0)


def _c1_cases(ctx, orc, vessel_grid):
    for dfunc in (0, 1, 2):
        seeds, _ = orc.seed_uniform(orc.Rng(80), vessel_grid, 12)
        lab = orc.naive(vessel_grid.copy(), seeds, dfunc)
        assert np.array_equal(_run_c1(ctx, lab, seeds), orc.remove_isolated_regions_cpu(lab.copy(), seeds))
        dense = np.ones((96, 80, 128), np.uint16)
        ds = pick_seeds(dense, 40, dfunc)
        dl = orc.naive(dense.copy(), ds, dfunc)
        assert np.array_equal(_run_c1(ctx, dl, ds), orc.remove_isolated_regions_cpu(dl.copy(), ds))
        assert np.array_equal(_run_c1(ctx, dl, ds[:30]), orc.remove_isolated_regions_cpu(dl.copy(), ds[:30]))  # ten labels lose their seed
    for trial in range(5):
        g = random_blob_grid((35, 33, 72 if trial % 2 else 70), 300 + trial, fill=0.45, smooth=1)
        seeds = pick_seeds(g, 8, trial)
        lab = orc.naive(g.copy(), seeds, 1)
        lab[1:3, 1:3, 1:3] = 1
        s2 = seeds.copy()
        s2[0, :3] = np.argwhere(lab == s2[1, 3])[0]
        for sd in (seeds, s2):
            assert np.array_equal(_run_c1(ctx, lab, sd), orc.remove_isolated_regions_cpu(lab.copy(), sd))


@pytest.mark.parametrize("limit,mode", [(0, 4), (1, 4), (40, 4), (40, 0), (700, 2), (8192, 4), (8192, 0), (65536, 4)])
def test_result_does_not_depend_on_the_front_limit(orc, vessel_grid, limit, mode):
    """setFloodFront: every phase starts level by level on one thread-block cluster and moves to the tile worklist once a level holds more than
    `limit` cells — never (65536 on these grids), at once (1: the seeds themselves are too many), somewhere on the way (40, 700), or not at all
    (0: tiles only).  Plain floods, extra seeds (both phases start on the front solver), a solid block (wide fronts), a labyrinth, ragged dims
    without TMA staging, a seed on an EMPTY cell, two seeds on one cell.  Labels, max_dist and the F3 counters stay bit-exact."""
    import voxelfragmentml_b200 as vf

    c = vf.Context(0)
    c.setFloodFront(limit)
    c.setFloodMode(mode)
    seeds, _ = orc.seed_uniform(orc.Rng(80), vessel_grid, 16)
    for dfunc in (1, 2):
        want, st = orc.flood(vessel_grid.copy(), seeds, dfunc)
        got, gst = _run_flood(c, vessel_grid, seeds, dfunc)
        assert np.array_equal(got, want) and gst.max_dist == st.max_dist
        if limit == 0:
            assert gst.front_levels == 0 and gst.tile_visits >= 1
        if limit == 65536:
            assert gst.front_levels >= st.max_dist + 1 and gst.tile_visits == 0  # the front solver ran every level, the tiles had nothing to do
        xs = orc.make_seeds(orc.Rng(85), vessel_grid, 5, 10)
        want, st = orc.flood(vessel_grid.copy(), xs, dfunc)
        got, gst = _run_flood(c, vessel_grid, xs, dfunc)
        assert np.array_equal(got, want) and gst.disjoint_rounds == st.rounds and gst.freed_voxels == st.freed_voxels
    solid = np.ones((48, 40, 64), np.uint16)
    solid[10:20, 5:30, 8:50] = 0
    sd = pick_seeds(solid, 7, 3)
    sd15 = sd.copy()
    sd15[:, 3] = np.uint32([300, 301, 2, 32767, 5000, 77, 1024])
    for dfunc in (1, 2):
        want, st = orc.flood(solid.copy(), sd, dfunc)
        got, gst = _run_flood(c, solid, sd, dfunc)
        assert np.array_equal(got, want) and gst.max_dist == st.max_dist
        want, st = orc.flood(solid.copy(), sd15, dfunc, id_bits=15)
        got, gst = _run_flood(c, solid, sd15, dfunc, id_bits=15)
        assert np.array_equal(got, want) and gst.max_dist == st.max_dist
    lab = np.zeros((48, 40, 64), np.uint16)  # serpentine corridor
    lab[1:-1, 1:-1, 1:-1] = 1
    for x in range(4, 44, 4):
        lab[x, (1 if (x // 4) % 2 else 4):(36 if (x // 4) % 2 else 39), :] = 0
    sd = np.uint32([[2, 2, 2, 2], [45, 37, 60, 3]])
    want, st = orc.flood(lab.copy(), sd, 1)
    got, gst = _run_flood(c, lab, sd, 1)
    assert np.array_equal(got, want) and gst.max_dist == st.max_dist
    rag = random_blob_grid((33, 21, 75), 9, fill=0.6)  # Z % 4 != 0
    sd = pick_seeds(rag, 5, 2)
    for dfunc in (1, 2):
        want, _ = orc.flood(rag.copy(), sd, dfunc)
        got, _ = _run_flood(c, rag, sd, dfunc)
        assert np.array_equal(got, want)
    if limit >= 8192:
        # a solid block: (a) one seed in the middle — its front stays with one CTA, which deals it out through the inboxes (2048 pairs per barrier
        # interval and CTA): the 26-neighbour front outgrows them long before 65536 pairs are pending, lowered cells go unlisted and the tiles have to
        # find them; (b) two corners; (c) 3000 seeds, whose first front is dealt over the whole cluster
        big = np.ones((112, 96, 128), np.uint16)
        big[40:60, :70, 30:90] = 0
        for sd in (np.uint32([[70, 48, 64, 2]]), np.uint32([[0, 0, 0, 2], [111, 95, 127, 3]]), pick_seeds(big, 3000, 5)):
            for dfunc in (1, 2):
                want, st = orc.flood(big.copy(), sd, dfunc, id_bits=15)
                got, gst = _run_flood(c, big, sd, dfunc, id_bits=15)
                print(f"limit {limit} seeds {len(sd)} dfunc {dfunc}: front steps {gst.front_levels}, tile visits {gst.tile_visits}, max_dist {gst.max_dist}")
                assert np.array_equal(got, want) and gst.max_dist == st.max_dist and gst.front_levels >= 1
    g = np.zeros((20, 20, 40), np.uint16)
    g[2:8, 2:8, 2:30] = 1
    g[12:18, 12:18, 5:35] = 1  # island without a seed stays FREE
    sd = np.array([[3, 3, 3, 2], [7, 7, 29, 3], [10, 10, 10, 4], [3, 3, 3, 5], [0, 0, 0, 6], [19, 19, 39, 7]], np.uint32)  # EMPTY cells, a shared cell, grid corners
    for dfunc in (1, 2):
        want, _ = orc.flood(g.copy(), sd, dfunc)
        got, _ = _run_flood(c, g, sd, dfunc)
        assert np.array_equal(got, want)
    c.close()


def test_front_limit_argument_is_checked():
    import voxelfragmentml_b200 as vf

    c = vf.Context(0)
    with pytest.raises(vf.VoxFragError):
        c.setFloodFront(65537)  # the start list holds 65536 cells
    c.setFloodFront(65536)
    c.close()
