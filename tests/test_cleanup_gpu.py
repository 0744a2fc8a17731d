"""C2-C4 / H1 / S1-S2 / X1 / fractureModel parity on the GPU against the oracle (bit-exact)."""
import ctypes as C
import os
import struct

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


@pytest.fixture(scope="module")
def labelled_vessel(orc, vessel_grid):
    seeds, _ = orc.seed_uniform(orc.Rng(80), vessel_grid, 8)
    return orc.naive(vessel_grid.copy(), seeds, 0), seeds


def _grid(ctx, host):
    import voxelfragmentml_b200 as vf

    g = vf.RegularGrid(ctx, host.shape)
    g.updateSSBO(host)
    return g


def test_detect_boundaries_twice_and_undo_mask(ctx, orc, labelled_vessel):
    lab, _ = labelled_vessel
    g = _grid(ctx, lab)
    g.detectBoundaries(1)
    want = orc.detect_boundaries(lab.copy(), 1)
    got = g.updateGrid()
    assert np.array_equal(got, want) and (got & 0x8000).any()
    g.detectBoundaries(1)  # erode re-applies it without undoMask (RegularGrid.cpp:133-135)
    assert np.array_equal(g.updateGrid(), orc.detect_boundaries(want.copy(), 1))
    g.undoMask()
    assert np.array_equal(g.updateGrid(), lab)
    g.close()


@pytest.mark.parametrize("shape", [(19, 23, 70), (8, 8, 64), (9, 17, 65), (3, 2, 5), (19, 23, 68), (16, 9, 140), (9, 9, 4), (8, 8, 12), (20, 20, 76)])
def test_detect_boundaries_ragged(ctx, orc, shape):
    g0 = random_blob_grid(shape, 1, fill=0.6, smooth=0)
    lab = orc.naive(g0.copy(), pick_seeds(g0, min(5, int((g0 != 0).sum())), 3), 1)
    g = _grid(ctx, lab)
    g.detectBoundaries(1)
    assert np.array_equal(g.updateGrid(), orc.detect_boundaries(lab.copy(), 1))
    g.close()


@pytest.mark.parametrize("etype,size,iters,mode", [(1, 3, 3, 0), (2, 3, 3, 0), (0, 3, 2, 0), (1, 3, 1, 1), (1, 3, 3, 1), (1, 5, 2, 0), (0, 4, 1, 0), (1, 3, 0, 0)])
def test_erode(ctx, orc, labelled_vessel, etype, size, iters, mode):
    lab, _ = labelled_vessel
    noise = orc.Rng(80).fill_noise(100003)
    want = orc.erode(lab.copy(), noise, etype, size, iters, 0.5, 0.5, boundary_mode=mode)
    g = _grid(ctx, lab)
    g.erode(etype, size, iters, 0.5, 0.5, noise=noise, boundaryMode=mode)
    got = g.updateGrid()
    assert np.array_equal(got, want)
    if iters and mode == 0:
        assert (got != lab).any()
    g.close()


def test_erode_probability_threshold_variants(ctx, orc):
    g0 = random_blob_grid((30, 26, 70), 8, fill=0.55, smooth=1)
    lab = orc.naive(g0.copy(), pick_seeds(g0, 6, 1), 0)
    noise = orc.Rng(5).fill_noise(4099)
    for prob, thr in [(0.1, 0.9), (1.0, 0.3), (0.7, 0.7)]:
        want = orc.erode(lab.copy(), noise, 1, 3, 3, prob, thr)
        g = _grid(ctx, lab)
        g.erode(1, 3, 3, prob, thr, noise=noise)
        assert np.array_equal(g.updateGrid(), want)
        g.close()


@pytest.mark.parametrize("shape", [(30, 26, 68), (17, 40, 140), (12, 12, 12), (9, 10, 4)])
def test_stencils_on_grids_whose_z_is_a_multiple_of_4_only(ctx, orc, shape):
    """Z % 8 == 4: the fast stencil path without the TMA box (the reference's dataset dims rule yields such grids, e.g. 140 x 200 x 140):
    erosion (all three masks), boundary tags in and out, the 3^3 sweep."""
    g0 = random_blob_grid(shape, 8, fill=0.6, smooth=1)
    lab = orc.naive(g0.copy(), pick_seeds(g0, min(6, int((g0 != 0).sum())), 1), 0)
    noise = orc.Rng(5).fill_noise(4099)
    for etype, mode in [(1, 0), (0, 0), (2, 1)]:
        want = orc.erode(lab.copy(), noise, etype, 3, 3, 0.6, 0.6, boundary_mode=mode)
        g = _grid(ctx, lab)
        g.erode(etype, 3, 3, 0.6, 0.6, noise=noise, boundaryMode=mode)
        assert np.array_equal(g.updateGrid(), want), (shape, etype, mode)
        g.close()
    speck = lab.copy()
    speck[::3, ::2, ::3] = 9
    g = _grid(ctx, speck)
    g.removeIsolatedRegions()
    assert np.array_equal(g.updateGrid(), orc.remove_isolated_regions_grid(speck.copy()))
    g.close()


def test_remove_isolated_regions_grid(ctx, orc, labelled_vessel):
    lab, _ = labelled_vessel
    speck = lab.copy()
    speck[::7, ::5, ::3] = 9  # isolated voxels of a foreign label
    g = _grid(ctx, speck)
    g.removeIsolatedRegions()
    assert np.array_equal(g.updateGrid(), orc.remove_isolated_regions_grid(speck.copy()))
    g.close()


def test_pointwise_passes(ctx, orc):
    rs = np.random.RandomState(0)
    host = rs.randint(0, 65536, size=(7, 9, 13)).astype(np.uint16)  # 819 cells: exercises the vector tail
    for name, fn in [("undoMask", lambda a: orc.undo_mask(a, 15, False)), ("resetFilling", orc.reset_filling), ("homogenize", orc.homogenize)]:
        g = _grid(ctx, host)
        getattr(g, name)()
        assert np.array_equal(g.updateGrid(), fn(host.copy())), name
        g.close()


def test_histogram(ctx, orc, labelled_vessel):
    lab, _ = labelled_vessel
    tagged = orc.detect_boundaries(lab.copy(), 1)
    for host in (lab, tagged):
        g = _grid(ctx, host)
        counts, occ = g.countValues()
        wc, wo = orc.count_values(host)
        assert occ == wo and np.array_equal(counts, wc)
        assert g.numOccupiedVoxels() == wo
        g.close()
    big = (np.random.RandomState(1).randint(0, 20000, size=(31, 17, 9)).astype(np.uint16))  # ids beyond the shared-memory bins
    g = _grid(ctx, big)
    counts, occ = g.countValues()
    wc, wo = orc.count_values(big)
    assert occ == wo and np.array_equal(counts, wc)
    g.close()


def test_rng_stream_and_noise(ctx, orc):
    ctx.initSeed(80)
    r = orc.Rng(80)
    assert [ctx.rng_raw() for _ in range(1000)] == [r.raw() for _ in range(1000)]
    assert [ctx.getUniformRandom() for _ in range(1000)] == [r.uniform() for _ in range(1000)]
    ctx.initSeed(7)
    assert np.array_equal(ctx.fillNoiseBuffer(5000), orc.Rng(7).fill_noise(5000))


@pytest.mark.parametrize("location", [0, 1, 2])
def test_seed_uniform_matches_reference_stream(ctx, orc, vessel_grid, location):
    import voxelfragmentml_b200 as vf

    g = _grid(ctx, vessel_grid)
    for n in (1, 8, 64):
        ctx.initSeed(80 + n)
        r = orc.Rng(80 + n)
        got, att = vf.Seeder.uniform(g, n, location=location, return_attempts=True)
        want, watt = orc.seed_uniform(r, vessel_grid, n, location=location)
        assert np.array_equal(got, want) and att == watt
        assert ctx.rng_raw() == r.raw()  # generator left exactly where the reference's would be
    g.close()


@pytest.mark.parametrize("location", [0, 1, 2])
def test_seed_halton_matches_reference_sampler(ctx, orc, vessel_grid, location):
    """S3: HALTON seeding (Faure-permuted sampler, pinned against the reference header in tests/test_oracle_vs_ref.py): same
    seeds, same attempt count, and the mt19937 state untouched (the Halton mode draws nothing from it)."""
    import voxelfragmentml_b200 as vf

    g = _grid(ctx, vessel_grid)
    for n in (1, 8, 64, 300):
        ctx.initSeed(7)
        before = ctx.rng_raw()
        ctx.initSeed(7)
        got, att = vf.Seeder.uniform(g, n, randomSeedFunction=vf.RandomUniformType.HALTON, location=location, return_attempts=True)
        want, watt = orc.seed_uniform(orc.Rng(7), vessel_grid, n, mode=1, location=location)
        assert np.array_equal(got, want) and att == watt
        assert ctx.rng_raw() == before
    with pytest.raises(vf.VoxFragError):
        vf.Seeder.uniform(g, 4, randomSeedFunction=vf.RandomUniformType.BOOST_NORMAL_DISTRIBUTION)
    g.close()


def test_make_seeds_with_extras_and_exhaustion(ctx, orc, vessel_grid):
    import voxelfragmentml_b200 as vf

    g = _grid(ctx, vessel_grid)
    for n, ne, md in [(8, 16, 0), (3, 6, 1), (10, 20, 2)]:
        ctx.initSeed(123)
        got = vf.Seeder.make(g, n, ne, mergeDFunc=md)
        want = orc.make_seeds(orc.Rng(123), vessel_grid, n, ne, merge_dfunc=md)
        assert np.array_equal(got, want)
    g.close()
    empty = vf.RegularGrid(ctx, (16, 16, 16))
    with pytest.raises(vf.SeederSearchError):
        vf.Seeder.uniform(empty, 1)
    empty.close()
    with pytest.raises(vf.VoxFragError) as e:
        vf.Seeder.uniform(_grid(ctx, vessel_grid), 2, randomSeedFunction=vf.RandomUniformType.BOOST_NORMAL_DISTRIBUTION)
    assert e.value.status == 7


def test_fracture_model_flood_defaults(ctx, orc, vessel_grid):
    """CADScene::fractureModel with the reference defaults: FLOOD + CHEBYSHEV, 8 seeds + 16 extra, detectBoundaries(1)."""
    import voxelfragmentml_b200 as vf

    p = vf.FractureParameters()
    g = _grid(ctx, vessel_grid)
    ctx.initSeed(p._seed)
    seeds, st = vf.fracture_model(g, p)
    r = orc.Rng(80)
    wseeds = orc.make_seeds(r, vessel_grid, 8, 16, merge_dfunc=0)
    assert np.array_equal(seeds, wseeds)
    want, wst = orc.flood(vessel_grid.copy(), wseeds, orc.CHEBYSHEV)
    want = orc.detect_boundaries(want, 1)
    assert np.array_equal(g.updateGrid(), want)
    assert st.disjoint_rounds == wst.rounds
    g.close()


def test_fracture_model_naive_erode(ctx, orc, vessel_grid):
    import voxelfragmentml_b200 as vf

    p = vf.FractureParameters(_fractureAlgorithm=vf.FractureAlgorithm.NAIVE, _distanceFunction=vf.DistanceFunction.EUCLIDEAN, _numSeeds=6,
                              _numExtraSeeds=0, _erode=1)
    g = _grid(ctx, vessel_grid)
    ctx.initSeed(80)
    seeds, _ = vf.fracture_model(g, p)
    r = orc.Rng(80)
    wseeds = orc.make_seeds(r, vessel_grid, 6, 0)
    want = orc.naive(vessel_grid.copy(), wseeds, 0)
    want = orc.remove_isolated_regions_cpu(want, wseeds)
    want = orc.erode(want, r.fill_noise(1000000), 1, 3, 3, 0.5, 0.5)
    assert np.array_equal(seeds, wseeds) and np.array_equal(g.updateGrid(), want)
    g.close()


def test_export_rle_and_bing(ctx, orc, labelled_vessel, tmp_path):
    import voxelfragmentml_b200 as vf

    lab, _ = labelled_vessel
    g = _grid(ctx, lab)
    base = str(tmp_path / "AL_12B_8f_128r_0it")
    g.exportGrid(base, True, vf.ExportGrid.RLE)
    assert open(base + ".rle", "rb").read() == orc.encode_rle(lab)
    g.exportGrid(base, True, vf.ExportGrid.UNCOMPRESSED_BINARY)
    assert open(base + ".bing", "rb").read() == orc.encode_bing_squared(lab)
    for squared in (False, True):
        g.exportGrid(base, squared, vf.ExportGrid.VOX)
        assert open(base + ".vox", "rb").read() == orc.encode_vox(lab, squared)
    g.exportGrid(base, True, vf.ExportGrid.QUADSTACK)
    assert open(base + ".qstack", "rb").read() == orc.encode_qstack(lab)
    g.close()


def test_device_rle_encoder_matches_oracle(ctx, orc, labelled_vessel):
    """vf_grid_encode_rle: runs found on the device == exportRLE's bytes, on label grids, noise (one run per cell), uniform
    grids (one run), cell counts that are not a multiple of the 8-cell vectors or of the 2048-cell tiles, and 1-cell grids."""
    lab, _ = labelled_vessel
    rs = np.random.RandomState(9)
    grids = [lab, np.ones((1, 1, 1), np.uint16), np.zeros((64, 64, 64), np.uint16), np.full((3, 5, 7), 4, np.uint16),
             rs.randint(0, 3, size=(17, 13, 11)).astype(np.uint16), rs.randint(0, 60000, size=(40, 41, 43)).astype(np.uint16),
             np.repeat(rs.randint(0, 9, size=(33, 9, 5)).astype(np.uint16), 13, axis=2)]
    for a in grids:
        g = _grid(ctx, a)
        want = orc.encode_rle(a)
        got = g.encodeRLE()
        assert got == want, a.shape
        # size query + too-small buffer leave the caller's memory alone
        need = C.c_uint64(0)
        small = np.zeros(8, np.uint8)
        assert g._lib.vf_grid_encode_rle(g._h, small.ctypes.data, 8, C.byref(need)) == 0
        assert need.value == len(want) and not small.any()
        g.close()


def test_near_seeds_and_fracture_model_with_impacts(ctx, orc, vessel_grid):
    """S3: Seeder::nearSeeds on the device-resident grid == the oracle (MSVC rand() LCG + mt19937), including both generators'
    positions afterwards; then CADScene::fractureModel with _numImpacts > 0 (uniform -> nearSeeds -> extra seeds -> flood)."""
    import voxelfragmentml_b200 as vf

    g = _grid(ctx, vessel_grid)
    frags = pick_seeds(vessel_grid, 6, 5)
    for impacts, nseeds, spreading, seed in [(1, 12, 5, 80), (3, 20, 3, 7), (2, 2, 8, 11)]:
        ctx.initSeed(seed)
        got = vf.Seeder.nearSeeds(g, frags, impacts, nseeds, spreading)
        r = orc.Rng(seed)
        want, st = orc.near_seeds(r, vessel_grid, frags, impacts, nseeds, spreading, crand_state=seed)
        assert np.array_equal(got, want)
        assert ctx._lib.vf_crand_next(ctx._h) == (((st * 214013 + 2531011) & 0xFFFFFFFF) >> 16) & 0x7FFF
        assert ctx.getUniformRandom() == r.uniform()
    p = vf.FractureParameters(_numSeeds=5, _numExtraSeeds=10, _numImpacts=2, _biasSeeds=9, _biasFocus=4)
    ctx.initSeed(80)
    seeds, _ = vf.fracture_model(g, p)
    r = orc.Rng(80)
    s0, _ = orc.seed_uniform(r, vessel_grid, 5)
    s1, _ = orc.near_seeds(r, vessel_grid, s0, 2, 9, 4, crand_state=80)
    extra, _ = orc.seed_uniform(r, vessel_grid, 10, location=orc.BOTH)
    merged = orc.merge_seeds(s1, np.concatenate([s1, extra]), 0)
    wseeds = np.concatenate([s1, merged])
    assert np.array_equal(seeds, wseeds)
    want, _ = orc.flood(vessel_grid.copy(), wseeds, orc.CHEBYSHEV)
    assert np.array_equal(g.updateGrid(), orc.detect_boundaries(want, 1))
    g.close()


def test_histogram_undo_mask_fused(ctx, orc, labelled_vessel):
    """vf_histogram_undo_mask == countValues followed by undoMask (prepareScene's grid side), incl. cell counts not divisible by 8"""
    lab, _ = labelled_vessel
    rs = np.random.RandomState(4)
    tagged = orc.detect_boundaries(lab.copy(), 1)
    odd = (rs.randint(0, 40, size=(13, 7, 9)) | (rs.rand(13, 7, 9) < 0.3) * 0x8000).astype(np.uint16)
    big_ids = (rs.randint(0, 30000, size=(16, 16, 40)) | (rs.rand(16, 16, 40) < 0.5) * 0x8000).astype(np.uint16)
    for a in (tagged, odd, big_ids, np.zeros((8, 8, 8), np.uint16)):
        g = _grid(ctx, a)
        counts, occ = g.countValuesUndoMask()
        wcounts, wocc = orc.count_values(a)
        assert occ == wocc and np.array_equal(counts, wcounts)
        assert np.array_equal(g.updateGrid(), orc.undo_mask(a.copy(), 15, False))
        g.close()


def test_connected_to_seed_with_whole_tile_regions(ctx, orc):
    """C1 on grids whose 16 x 16 x 32 tiles are mostly interior to one label (the union-find shortcut for whole-region tiles):
    solid blocks split by planes, an island of the same label cut off by another label, tiles cut by the grid border."""
    import voxelfragmentml_b200 as vf

    rs = np.random.RandomState(21)
    for dims in [(64, 48, 96), (40, 40, 72), (33, 50, 64)]:
        g0 = np.full(dims, 2, np.uint16)
        g0[dims[0] // 2:, :, :] = 3
        g0[:, dims[1] // 2:, dims[2] // 2:] = 4
        g0[4:12, 4:12, 4:20] = 3            # an island of label 3 inside label 2: not connected to seed 3's region
        g0[:, :, dims[2] - 5] = rs.randint(2, 5, size=dims[:2])  # a noisy plane
        g0[20:22, :, :] = 0                 # an EMPTY slab: label 2 and 4 regions beyond it survive only with their own seed
        seeds = np.uint32([[1, 1, 1, 2], [dims[0] - 2, 1, 1, 3], [1, dims[1] - 2, dims[2] - 2, 4]])
        g = _grid(ctx, g0)
        vf.NaiveFracturer.removeIsolatedRegions(g, seeds)
        assert np.array_equal(g.updateGrid(), orc.remove_isolated_regions_cpu(g0.copy(), seeds)), dims
        g.close()


@pytest.mark.parametrize("shape", [(128, 110, 128), (17, 19, 35), (9, 5, 13), (40, 36, 64)])
def test_upload_bits_expands_one_bit_per_cell(ctx, orc, vessel_grid, shape):
    """vf_grid_upload_bits: bit (i & 7) of byte i >> 3 is cell i; set -> FREE, clear -> EMPTY (cell counts that are not multiples of 8 included)"""
    import voxelfragmentml_b200 as vf

    occ = vessel_grid if shape == vessel_grid.shape else random_blob_grid(shape, 3, fill=0.5, smooth=0)
    g = vf.RegularGrid(ctx, shape)
    g.fillValue(7)
    bits = np.packbits((occ.reshape(-1) != 0).astype(np.uint8), bitorder="little")
    g.upload_bits(bits)
    assert np.array_equal(g.updateGrid(), (occ != 0).astype(np.uint16))
    g.close()


@pytest.mark.parametrize("b", [0, 2, 3, 5])
def test_detect_boundaries_with_larger_boundary_size(ctx, orc, labelled_vessel, b):
    """detectBoundaries-comp.glsl:24-25 takes any boundarySize (the reference passes 1); the oracle's rule is pinned by the shader text itself
    (tests/test_oracle_vs_glsl.py).  Also on an already tagged grid and on a ragged one."""
    lab, _ = labelled_vessel
    for host in (lab, orc.detect_boundaries(lab.copy(), 1)):
        g = _grid(ctx, host)
        g.detectBoundaries(b)
        assert np.array_equal(g.updateGrid(), orc.detect_boundaries(host.copy(), b))
        g.close()
    g0 = random_blob_grid((19, 23, 37), 1, fill=0.6, smooth=0)
    rag = orc.naive(g0.copy(), pick_seeds(g0, 5, 3), 1)
    g = _grid(ctx, rag)
    g.detectBoundaries(b)
    assert np.array_equal(g.updateGrid(), orc.detect_boundaries(rag.copy(), b))
    g.close()


def test_qstack_export_from_device_runs(ctx, orc, tmp_path):
    """`.qstack` with the z-column runs taken from the device-side run stream (no full-grid download): every case whose bytes the reference's
    own QuadStack.h / GStack.h pin (tests/golden/qstack_golden.json through the oracle, tests/test_oracle_golden.py)"""
    import voxelfragmentml_b200 as vf
    from vox_cases import all_qstack_cases

    for name, grid in all_qstack_cases():
        g = _grid(ctx, grid)
        base = str(tmp_path / f"q_{name}")
        g.exportGrid(base, True, vf.ExportGrid.QUADSTACK)
        assert open(base + ".qstack", "rb").read() == orc.encode_qstack(grid), name
        g.close()


@pytest.mark.parametrize("iters", [2, 3])
def test_erode_when_passes_empty_millions_of_cells(ctx, orc, iters):
    """threshold and probability high enough that a pass empties far more than the 2^20 cells the change lists hold: the sparse erosion passes,
    the sparse sweep and its apply step then scan the change bitmap instead (odd and even iteration counts: the caller's grid holds the output
    of the last pass or of the one before)"""
    g0 = np.ones((160, 128, 192), np.uint16)
    lab = orc.naive(g0.copy(), pick_seeds(g0, 400, 4), 0)  # small regions: most cells sit next to a border
    noise = orc.Rng(11).fill_noise(100003)
    orc.use_all_cores()
    want = orc.erode(lab.copy(), noise, 1, 3, iters, 0.5, 1.6)  # every cell whose noise value is below 0.5 goes in the first pass
    assert int(((want == 0) & (lab != 0)).sum()) > (1 << 20) and (want != 0).any()
    g = _grid(ctx, lab)
    g.erode(1, 3, iters, 0.5, 1.6, noise=noise)
    assert np.array_equal(g.updateGrid(), want)
    g.close()
