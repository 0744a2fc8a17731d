"""Pins the oracle against the REFERENCE'S OWN CODE compiled in place (oracle/_ref/libvf_ref.so, built by oracle/ref_shim/Makefile
from /root/reference: NaiveFracturer.cpp buildCPU + removeIsolatedRegionsCPU, Seeder.cpp uniform + mergeSeeds with the real
RandomUtilities.h, the Möller SAT of Intersections3D.h:204-420, AABB.cpp).  Skipped when the library has not been built."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ROOT, pick_seeds, random_blob_grid
from vox_cases import all_qstack_cases, vox_cases

REF_SO = os.path.join(ROOT, "oracle", "_ref", "libvf_ref.so")
pytestmark = pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref not built (needs /root/reference)")

_u16 = np.ctypeslib.ndpointer(np.uint16, flags="C_CONTIGUOUS")
_u32 = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_f32 = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")


@pytest.fixture(scope="module")
def ref():
    L = C.CDLL(REF_SO)
    L.ref_naive_build_cpu.argtypes = [_u16, _u32, _u32, C.c_uint32, C.c_int]
    L.ref_remove_isolated_regions_cpu.argtypes = [_u16, _u32, _u32, C.c_uint32]
    L.ref_seed_uniform.restype = C.c_int
    L.ref_seed_uniform.argtypes = [_u16, _u32, C.c_uint32, C.c_int, C.c_int, _u32, C.POINTER(C.c_float)]
    L.ref_merge_seeds.argtypes = [_u32, C.c_uint32, _u32, C.c_uint32, C.c_int]
    L.ref_tri_box_intersect.restype = C.c_int
    L.ref_tri_box_intersect.argtypes = [_f32] * 5
    L.ref_uniform_draws.argtypes = [C.c_int, C.c_int, _f32]
    return L


def _dims(g):
    return np.asarray(g.shape, np.uint32)


def test_rng_stream_is_the_reference_stream(ref, orc):
    """RandomUtilities (std::mt19937 + uniform_real_distribution<float>) under this toolchain == the hard-coded recipe."""
    for seed in (80, 1, 2024):
        out = np.zeros(20000, np.float32)
        ref.ref_uniform_draws(seed, len(out), out)
        r = orc.Rng(seed)
        mine = np.array([r.uniform() for _ in range(len(out))], np.float32)
        assert np.array_equal(out.view(np.uint32), mine.view(np.uint32))
    assert np.allclose(out[:0], [])  # keep flake quiet


@pytest.mark.parametrize("dfunc", [0, 1, 2])
def test_naive_buildcpu(ref, orc, vessel_grid, dfunc):
    for grid, ns in [(vessel_grid, 8), (random_blob_grid((31, 22, 40), 5) * 3, 17)]:
        grid = grid.astype(np.uint16)
        seeds = pick_seeds(grid, ns, ns)
        a = grid.copy()
        ref.ref_naive_build_cpu(a, _dims(a), seeds, len(seeds), dfunc)
        b = orc.naive(grid.copy(), seeds, dfunc)
        assert np.array_equal(a, b)


def test_remove_isolated_regions_cpu(ref, orc, vessel_grid):
    for trial in range(4):
        g = random_blob_grid((33, 30, 41), 50 + trial, fill=0.45, smooth=1)
        seeds = pick_seeds(g, 8, trial)
        lab = orc.naive(g.copy(), seeds, 1)
        a = lab.copy()
        ref.ref_remove_isolated_regions_cpu(a, _dims(a), seeds, len(seeds))
        b = orc.remove_isolated_regions_cpu(lab.copy(), seeds)
        assert np.array_equal(a, b) and (a != lab).any()
    seeds, _ = orc.seed_uniform(orc.Rng(80), vessel_grid, 12)
    lab = orc.naive(vessel_grid.copy(), seeds, 0)
    a = lab.copy()
    ref.ref_remove_isolated_regions_cpu(a, _dims(a), seeds, len(seeds))
    assert np.array_equal(a, orc.remove_isolated_regions_cpu(lab.copy(), seeds))


@pytest.mark.parametrize("location", [0, 1, 2])
def test_seeder_uniform(ref, orc, vessel_grid, location):
    for n, seed in [(8, 80), (32, 81), (1, 5)]:
        out = np.zeros((n, 4), np.uint32)
        nxt = C.c_float(0)
        g = vessel_grid.copy()
        assert ref.ref_seed_uniform(g, _dims(g), n, location, seed, out, C.byref(nxt)) == 0
        r = orc.Rng(seed)
        mine, _ = orc.seed_uniform(r, vessel_grid, n, location=location)
        assert np.array_equal(out, mine)
        assert np.float32(nxt.value) == np.float32(r.uniform())  # both generators stand at the same draw
    empty = np.zeros((8, 8, 8), np.uint16)
    assert ref.ref_seed_uniform(empty, _dims(empty), 1, 1, 1, np.zeros((1, 4), np.uint32), None) == -1  # SeederSearchError


def test_halton_sampler_and_seeding(ref, orc, vessel_grid):
    """S3: the Faure-permuted Halton sampler (dimensions 0..2) bit for bit, then Seeder::uniform in HALTON mode on the vessel."""
    ref.ref_halton.restype = C.c_float
    ref.ref_halton.argtypes = [C.c_uint, C.c_uint]
    rs = np.random.RandomState(5)
    idx = np.concatenate([np.arange(0, 5000), rs.randint(0, 2 ** 31 - 1, 5000), [999999, 1000000, 3 ** 20 - 1, 5 ** 12 - 1, 2 ** 32 - 1]])
    for d in (0, 1, 2):
        for i in idx:
            a, b = ref.ref_halton(d, int(i)), orc.halton(d, int(i))
            assert np.float32(a).tobytes() == np.float32(b).tobytes(), (d, int(i), a, b)
    ref.ref_seed_halton.restype = C.c_int
    ref.ref_seed_halton.argtypes = [_u16, _u32, C.c_uint32, C.c_int, _u32]
    grids = [vessel_grid, (random_blob_grid((37, 50, 29), 3) * 1).astype(np.uint16)]
    for g in grids:
        g = np.ascontiguousarray(g)
        for location in (0, 1, 2):
            for n in (1, 8, 40):
                a = np.zeros((n, 4), np.uint32)
                rc = ref.ref_seed_halton(g.copy(), _dims(g), n, location, a)
                try:
                    b, _ = orc.seed_uniform(orc.Rng(80), g, n, mode=1, location=location)
                except orc.OracleError:
                    assert rc == -1
                    continue
                assert rc == 0 and np.array_equal(a, b), (g.shape, location, n)


def test_merge_seeds(ref, orc):
    rs = np.random.RandomState(9)
    for dfunc in (0, 1, 2):
        frags = np.concatenate([rs.randint(0, 100, size=(9, 3)), np.arange(2, 11)[:, None]], 1).astype(np.uint32)
        extra = np.concatenate([rs.randint(0, 100, size=(30, 3)), np.zeros((30, 1), int)], 1).astype(np.uint32)
        seeds = np.ascontiguousarray(np.concatenate([frags, extra]))
        a = seeds.copy()
        ref.ref_merge_seeds(frags, len(frags), a, len(a), dfunc)
        assert np.array_equal(a, orc.merge_seeds(frags, seeds, dfunc))


def test_sat_predicate_matches_reference_bit_for_bit(ref, orc):
    """2e5 random triangle/box pairs, many of them grazing: identical verdicts (same float32 operation order)."""
    rs = np.random.RandomState(3)
    hits = 0
    for i in range(200000):
        c = rs.uniform(-0.5, 0.5, 3).astype(np.float32)
        h = np.float32(rs.choice([1 / 64, 1 / 128, 1 / 32]))
        bmin, bmax = (c - h).astype(np.float32), (c + h).astype(np.float32)
        scale = rs.choice([0.02, 0.05, 0.3])
        tri = (c + rs.normal(0, scale, size=(3, 3))).astype(np.float32)
        if i % 3 == 0:  # vertex exactly on a face / edge of the box
            tri[0] = bmin + (bmax - bmin) * rs.randint(0, 2, 3).astype(np.float32)
        a = ref.ref_tri_box_intersect(tri[0].copy(), tri[1].copy(), tri[2].copy(), bmin, bmax)
        b = orc.tri_box_intersect(tri[0], tri[1], tri[2], bmin, bmax)
        assert a == int(b), (i, tri, bmin, bmax)
        hits += a
    assert 20000 < hits < 180000


def _ref_vox(ref, grid, squared, tmp_path):
    ref.ref_export_vox.argtypes = [_u16, _u32, C.c_int, C.c_char_p]
    ref.ref_export_vox.restype = None
    path = str(tmp_path / "ref.vox")
    if os.path.exists(path):
        os.remove(path)
    ref.ref_export_vox(np.ascontiguousarray(grid), _dims(grid), int(squared), path.encode())
    return open(path, "rb").read()


@pytest.mark.parametrize("name,grid", vox_cases(), ids=[c[0] for c in vox_cases()])
@pytest.mark.parametrize("squared", [False, True])
def test_vox_bytes_match_reference_writer(ref, orc, tmp_path, name, grid, squared):
    """orc_encode_vox and the product's vf_encode_vox (host code) == the bytes the reference's VoxWriter.cpp writes."""
    import voxelfragmentml_b200 as vf

    want = _ref_vox(ref, grid, squared, tmp_path)
    assert want[:4] == b"VOX " and len(want) >= 20
    got = orc.encode_vox(grid, squared)
    assert got == want
    lib = vf._capi.load()
    d = _dims(grid)
    need = lib.vf_encode_vox(grid.ctypes.data, d.ctypes.data, int(squared), None, 0)
    assert need == len(want)
    buf = np.zeros(need, np.uint8)
    assert lib.vf_encode_vox(grid.ctypes.data, d.ctypes.data, int(squared), buf.ctypes.data, need) == need
    assert buf.tobytes() == want


QSTACK_SO = os.path.join(ROOT, "oracle", "_ref", "libvf_ref_qstack.so")


@pytest.mark.skipif(not os.path.exists(QSTACK_SO), reason="oracle/_ref qstack bridge not built")
@pytest.mark.parametrize("name,grid", all_qstack_cases(), ids=[c[0] for c in all_qstack_cases()])
def test_qstack_bytes_match_reference_quadstack(orc, tmp_path, name, grid):
    """orc_encode_qstack and the product's vf_encode_qstack (host code) == the bytes the reference's QuadStack.h / GStack.h write
    for exportQuadStack's call sequence."""
    import sys

    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from make_qstack_golden import ref_qstack_bytes

    import voxelfragmentml_b200 as vf

    want = ref_qstack_bytes(grid, str(tmp_path))
    assert orc.encode_qstack(grid) == want
    lib = vf._capi.load()
    d = _dims(grid)
    need = lib.vf_encode_qstack(grid.ctypes.data, d.ctypes.data, None, 0)
    assert need == len(want)
    buf = np.zeros(need, np.uint8)
    assert lib.vf_encode_qstack(grid.ctypes.data, d.ctypes.data, buf.ctypes.data, need) == need
    assert buf.tobytes() == want


@pytest.mark.skipif(not os.path.exists(QSTACK_SO), reason="oracle/_ref qstack bridge not built")
def test_qstack_random_grids_match_reference_quadstack(orc, tmp_path):
    import sys

    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from make_qstack_golden import ref_qstack_bytes

    rs = np.random.RandomState(5)
    for i in range(30):
        shape = tuple(int(v) for v in rs.randint(1, 18, 3))
        g = (rs.randint(0, 2 + i % 4, size=shape) * (rs.rand(*shape) < 0.7)).astype(np.uint16)
        if i % 3 == 0:  # columns that repeat across (x, y): uniform regions and merges
            g[:] = rs.randint(0, 3, size=(1, 1, shape[2]))
            g[shape[0] // 2:, :, ::2] = 4
        assert orc.encode_qstack(g) == ref_qstack_bytes(g, str(tmp_path)), (i, shape)


def test_near_seeds_matches_reference_seeder(ref, orc, vessel_grid):
    """Seeder::nearSeeds compiled in place vs the oracle in crand_mode 1 (both consume this process's ::rand() after srand(seed) and
    RandomUtilities' generator after initSeed): identical seed lists, so everything but the C runtime's rand() itself is pinned."""
    libc = C.CDLL(None)
    ref.ref_near_seeds.restype = C.c_int
    ref.ref_near_seeds.argtypes = [_u16, _u32, _u32, C.c_uint32, C.c_uint, C.c_uint, C.c_uint, C.c_int, C.c_uint, _u32]
    for grid, nfr in [(vessel_grid.astype(np.uint16), 6), ((random_blob_grid((40, 36, 44), 3) != 0).astype(np.uint16), 4)]:
        frags = pick_seeds(grid, nfr, 5)
        for impacts, nseeds, spreading, seed in [(1, 12, 5, 80), (3, 20, 3, 7), (2, 2, 8, 11)]:
            out = np.zeros((nfr + nseeds, 4), np.uint32)
            n = ref.ref_near_seeds(grid.copy(), _dims(grid), frags, nfr, impacts, nseeds, spreading, seed, seed, out)
            libc.srand(seed)
            mine, _ = orc.near_seeds(orc.Rng(seed), grid, frags, impacts, nseeds, spreading, crand_mode=1)
            assert n == len(mine) and np.array_equal(out[:n], mine), (impacts, nseeds, spreading)
            assert n > nfr
