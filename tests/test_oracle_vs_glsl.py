"""Second pin of the oracle for the stages that exist only as GLSL in the reference: the reference's OWN compute shaders, compiled as C++
from the text where it lies under /root/reference (oracle/ref_shim/glsl2cpp.py + ref_glsl.cpp -> oracle/_ref/libvf_ref_glsl.so), driven by
restatements of the host loops (RegularGrid.cpp:64-159, 488-503, 1006-1015; FloodFracturer.cpp:98-191; NaiveFracturer.cpp:71-109).

Order-independent shaders (detectBoundaries, erodeGrid, copyGrid, undoMask, naiveFracturer, disjointSet) must match the oracle bit for bit;
marchingCubes + computeMortonCodes (f2) hand out vertex slots with an atomic counter: their triangle soup is compared as a set.
removeIsolatedRegionsGrid races with itself in place: it is compared under the "all reads before all writes" schedule, which is the snapshot
rule the oracle and the CUDA path adopt.  floodFracturer races by design (whichever invocation stores first claims a cell): it runs under one
legal schedule (ascending invocation index) and is compared through what every schedule must produce — the set of claimed cells, the BFS
level of every cell, "each cell's label comes from a neighbour one level closer" — plus the share of cells whose label equals the oracle's
lowest-seed-index rule.  Skipped when the library has not been built (needs /root/reference)."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ROOT, pick_seeds, random_blob_grid

SO = os.path.join(ROOT, "oracle", "_ref", "libvf_ref_glsl.so")
pytestmark = pytest.mark.skipif(not os.path.exists(SO), reason="oracle/_ref/libvf_ref_glsl.so not built (needs /root/reference)")

_u16 = np.ctypeslib.ndpointer(np.uint16, flags="C_CONTIGUOUS")
_u32 = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_f32 = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")


@pytest.fixture(scope="module")
def glsl():
    L = C.CDLL(SO)
    L.glsl_detect_boundaries.argtypes = [_u16, _u32, C.c_int]
    L.glsl_remove_isolated_regions_grid.argtypes = [_u16, _u32, C.c_int]
    L.glsl_undo_mask.argtypes = [_u16, _u32, C.c_uint32, C.c_int]
    L.glsl_erode.argtypes = [_u16, _u32, C.c_int, C.c_uint32, C.c_uint32, C.c_float, C.c_float, _f32, C.c_uint32, C.c_int]
    L.glsl_naive.argtypes = [_u16, _u32, _u32, C.c_uint32, C.c_int]
    L.glsl_marching_cubes.restype = C.c_int
    L.glsl_marching_cubes.argtypes = [_u16, _u32, C.c_uint32, _f32, _f32, C.c_uint32, C.c_float, C.c_uint32, C.c_float, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, _u32]
    L.glsl_mc_soup.restype = C.c_uint32
    L.glsl_mc_soup.argtypes = [_u16, _u32, C.c_uint32, _f32, _u32, C.c_uint32]
    L.glsl_flood.restype = C.c_int
    L.glsl_flood.argtypes = [_u16, _u32, _u32, C.c_uint32, C.c_int, C.c_uint32, _u32]
    return L


def _dims(g):
    return np.asarray(g.shape, np.uint32)


def _labelled_cases(orc, vessel_grid):
    """label grids with fragment borders, holes and grid-face contact"""
    out = []
    seeds, _ = orc.seed_uniform(orc.Rng(80), vessel_grid, 12)
    out.append(orc.naive(vessel_grid.copy(), seeds, 0))
    dense = np.ones((24, 20, 40), np.uint16)
    out.append(orc.naive(dense.copy(), pick_seeds(dense, 9, 3), 2))
    blob = random_blob_grid((23, 18, 31), 7, fill=0.6)
    out.append(orc.naive(blob.copy(), pick_seeds(blob, 6, 1), 1))
    return out


def test_detect_boundaries_shader(glsl, orc, vessel_grid):
    for lab in _labelled_cases(orc, vessel_grid):
        for b in (1, 2, 3):
            a = lab.copy()
            glsl.glsl_detect_boundaries(a, _dims(a), b)
            assert np.array_equal(a, orc.detect_boundaries(lab.copy(), b)), f"boundarySize {b}"
        # a second call on the tagged grid (RegularGrid::erode calls it once per iteration)
        a2 = a.copy()
        glsl.glsl_detect_boundaries(a2, _dims(a2), 1)
        assert np.array_equal(a2, orc.detect_boundaries(a.copy(), 1))


def test_undo_mask_shader(glsl, orc):
    g = np.random.RandomState(5).randint(0, 65536, size=(9, 7, 11)).astype(np.uint16)
    for pos, rightmost in ((15, 0), (8, 1)):
        a = g.copy()
        glsl.glsl_undo_mask(a, _dims(a), pos, rightmost)
        assert np.array_equal(a, orc.undo_mask(g.copy(), pos, bool(rightmost)))


@pytest.mark.parametrize("dfunc", [0, 1, 2])
def test_naive_shader(glsl, orc, vessel_grid, dfunc):
    for grid, ns in [(vessel_grid, 8), (random_blob_grid((31, 22, 40), 5) * 3, 17)]:
        grid = grid.astype(np.uint16)
        seeds = pick_seeds(grid, ns, ns)
        a = grid.copy()
        glsl.glsl_naive(a, _dims(a), seeds, len(seeds), dfunc)
        # decode_mode 1: the shaders' float index decode (voxel.glsl:6-14); identical to the exact decode below 2^24 cells
        assert np.array_equal(a, orc.naive(grid.copy(), seeds, dfunc, decode_mode=1))
        assert np.array_equal(a, orc.naive(grid.copy(), seeds, dfunc))


def test_sweep_shader_under_the_snapshot_schedule(glsl, orc, vessel_grid):
    for lab in _labelled_cases(orc, vessel_grid):
        lab = orc.detect_boundaries(lab.copy())
        lab[::5, ::3, ::4] = 0  # stray holes: cells with few equal neighbours
        a = lab.copy()
        glsl.glsl_remove_isolated_regions_grid(a, _dims(a), 1)
        assert np.array_equal(a, orc.remove_isolated_regions_grid(lab.copy()))
        # the in-place ascending schedule is another legal outcome: it can only differ where a cell's neighbours were emptied before it ran
        b = lab.copy()
        glsl.glsl_remove_isolated_regions_grid(b, _dims(b), 0)
        assert np.all((b == a) | (b == 0))


@pytest.mark.parametrize("etype", [0, 1, 2])
@pytest.mark.parametrize("size,iters", [(3, 3), (3, 1), (5, 2), (4, 1)])
def test_erode_shader_loop(glsl, orc, vessel_grid, etype, size, iters):
    noise = orc.Rng(1080).fill_noise(5000)
    for ci, lab in enumerate(_labelled_cases(orc, vessel_grid)):
        if size > 3 and ci == 0:
            continue  # the 5^3 mask over the 1.8 M-cell vessel grid adds nothing the small grids do not cover
        for prob, thr in ((0.5, 0.5), (0.9, 0.8)):
            a = lab.copy()
            glsl.glsl_erode(a, _dims(a), etype, size, iters, prob, thr, noise, len(noise), 1)
            want = orc.erode(lab.copy(), noise, etype, size, iters, prob, thr, boundary_mode=0)
            assert np.array_equal(a, want), f"type {etype} size {size} iters {iters} p {prob} thr {thr}: {int((a != want).sum())} cells differ"


def _neighbour_has(lab, level, nneigh):
    """for every cell: does a flood neighbour hold the same label one level closer?"""
    X, Y, Z = lab.shape
    ok = np.zeros(lab.shape, bool)
    pl = np.pad(lab.astype(np.int64), 1, constant_values=-1)
    pv = np.pad(level.astype(np.int64), 1, constant_values=-10)
    for dx in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dz in (-1, 0, 1):
                if (dx, dy, dz) == (0, 0, 0) or (nneigh == 6 and abs(dx) + abs(dy) + abs(dz) != 1):
                    continue
                nl = pl[1 + dx : 1 + dx + X, 1 + dy : 1 + dy + Y, 1 + dz : 1 + dz + Z]
                nv = pv[1 + dx : 1 + dx + X, 1 + dy : 1 + dy + Y, 1 + dz : 1 + dz + Z]
                ok |= (nl == lab) & (nv == level - 1)
    return ok


@pytest.mark.parametrize("dfunc", [1, 2])
def test_flood_shader_loop_without_extra_seeds(glsl, orc, vessel_grid, dfunc):
    for grid, ns in [(vessel_grid, 12), (random_blob_grid((28, 26, 33), 11, fill=0.5, smooth=1), 7)]:
        grid = grid.astype(np.uint16)
        seeds = pick_seeds(grid, ns, 40 + ns)
        a = grid.copy()
        stats = np.zeros(4, np.uint32)
        assert glsl.glsl_flood(a, _dims(a), seeds, len(seeds), dfunc, 8 * a.size, stats) == 0
        want, st = orc.flood(grid.copy(), seeds, dfunc)
        keys = orc.flood_keys(grid.copy(), seeds, dfunc)
        level = (keys >> 15).astype(np.int64)
        claimed = a > 1
        assert np.array_equal(claimed, want > 1) and np.array_equal(a == 1, want == 1) and np.array_equal(a == 0, want == 0)
        # every claimed cell that is not a seed cell got its label from a neighbour one BFS level closer
        is_seed = np.zeros(a.shape, bool)
        is_seed[seeds[:, 0], seeds[:, 1], seeds[:, 2]] = True
        assert np.all(_neighbour_has(a, np.where(claimed, level, -5), 6 if dfunc == 1 else 26)[claimed & ~is_seed])
        assert int(stats[0]) == int(level[claimed].max()) + 1  # one dispatch per BFS level, plus the last one that claims nothing
        assert int(stats[1]) == 1 and int(stats[2]) == 0
        # the racy claims and the lowest-seed-index rule agree except on ties between fronts
        assert float((a == want).mean()) > 0.97


def _components_ok(lab, seeds_principal, nneigh):
    """each fragment id is one connected region (flood neighbourhood) that holds its principal seed"""
    from scipy import ndimage

    st = ndimage.generate_binary_structure(3, 1 if nneigh == 6 else 3)
    for fid, cell in seeds_principal.items():
        comp, n = ndimage.label(lab == fid, structure=st)
        if n != 1 or comp[tuple(cell)] != 1:
            return False
    return True


@pytest.mark.parametrize("dfunc", [1, 2])
def test_flood_shader_loop_with_extra_seeds(glsl, orc, vessel_grid, dfunc):
    """F3: prefix merge in the flood, disjointSet / disjointSetStack, re-flood until nothing is freed, then the low byte (FloodFracturer.cpp:135-186)."""
    for grid, nf in [(vessel_grid.astype(np.uint16), 5), (random_blob_grid((30, 24, 36), 3, fill=0.62).astype(np.uint16), 4)]:
        seeds = orc.make_seeds(orc.Rng(80 + nf), grid, nf, 2 * nf)
        a = grid.copy()
        stats = np.zeros(4, np.uint32)
        assert glsl.glsl_flood(a, _dims(a), seeds, len(seeds), dfunc, 16 * a.size, stats) == 0
        want, st = orc.flood(grid.copy(), seeds, dfunc)
        assert np.array_equal(a > 1, want > 1) and np.array_equal(a == 1, want == 1)
        assert int(a.max()) <= 1 + nf and set(np.unique(a[a > 1])) == set(np.unique(want[want > 1]))
        principal = {int(s[3]) & 0xFF: s[:3] for s in seeds[:nf]}  # the original seeds come first (CADScene.cpp:647-655)
        nneigh = 6 if dfunc == 1 else 26
        assert _components_ok(a, principal, nneigh) and _components_ok(want, principal, nneigh)
        assert int(stats[1]) >= 1
        assert float((a == want).mean()) > 0.9


def test_marching_cubes_and_morton_shaders(glsl, orc, vessel_grid):
    """f2, the first two dispatches of MarchingCubes::triangulateFieldGPU (MarchingCubes.cpp:364-388, 445-453): the reference's
    marchingCubes-comp.glsl over the padded grid and computeMortonCodes-comp.glsl over its vertices give the oracle's triangle soup —
    the same triangles (bit-identical float32 vertices, boundary flags and Morton codes), in whatever order the atomic counter dealt."""
    for lab in _labelled_cases(orc, vessel_grid):
        tagged = orc.detect_boundaries(lab.copy())
        labels = [int(v) for v in np.unique(lab) if v > 1]
        for target in labels[:3] + labels[-1:]:
            pd = np.asarray(tagged.shape, np.uint32) + 2
            padded = np.ones(tuple(int(v) for v in pd), np.uint16)  # MarchingCubes::setGrid (:523-534): a ring of VOXEL_FREE around the grid
            padded[1:-1, 1:-1, 1:-1] = tagged
            wv, wm = orc.mc_soup(tagged, target)
            cap = len(wv) + 300
            gv, gm = np.zeros((cap, 4), np.float32), np.zeros(cap, np.uint32)
            n = glsl.glsl_mc_soup(padded, pd, target, gv.reshape(-1), gm, cap)
            assert n == len(wv) and n % 3 == 0 and n > 0
            # one row per triangle: 3 x (x, y, z, flag) as bit patterns + 3 codes; the vertex order inside a triangle is part of the result
            def rows(v, m):
                r = np.concatenate([v[:n].view(np.uint32).reshape(n // 3, 12), m[:n].reshape(n // 3, 3)], axis=1)
                return r[np.lexsort(r.T[::-1])]
            assert np.array_equal(rows(gv, gm), rows(wv, wm)), f"target {target}"
            assert gv[:n, 3].max() <= 1.0 and (gv[:n, 3] == 1.0).any()  # tagged cells give flagged triangles


@pytest.mark.parametrize("iters", [None, 2])
def test_marching_cubes_pipeline_shaders(glsl, orc, vessel_grid, iters):
    """f2 end to end (MarchingCubes::triangulateFieldGPU, MarchingCubes.cpp:364-407): march, Morton codes, sort, findSameVertices_01 / _02,
    buildMarchingCubesFaces, markBoundaryTriangles and both smoothSurface passes, every shader compiled in place and dispatched in ascending
    invocation order.  The oracle orders equal Morton codes by position to be schedule independent; where no two DIFFERENT positions share a
    code (checked below: true for grids of this size) that is the order the shaders produce under this schedule, so vertices (float32 bits)
    and faces must be identical, numbering included."""
    mn, mx = np.float32([-0.4, 0.1, -0.3]), np.float32([0.6, 1.3, 0.45])
    for lab in _labelled_cases(orc, vessel_grid)[1:]:
        tagged = orc.detect_boundaries(lab.copy())
        labels = [int(v) for v in np.unique(lab) if v > 1]
        for target in labels[:2] + labels[-1:]:
            sv, sm = orc.mc_soup(tagged, target)
            order = np.lexsort((sv[:, 2], sv[:, 1], sv[:, 0], sm))
            same_code = sm[order][1:] == sm[order][:-1]
            same_pos = (sv[order][1:, :3] == sv[order][:-1, :3]).all(axis=1)
            assert not (same_code & ~same_pos).any()  # premise of the comparison
            pd = np.asarray(tagged.shape, np.uint32) + 2
            padded = np.ones(tuple(int(v) for v in pd), np.uint16)
            padded[1:-1, 1:-1, 1:-1] = tagged
            d = np.asarray(tagged.shape, np.uint32)
            it = int(np.float32(max(tagged.shape)) * np.float32(0.048)) if iters is None else iters
            kw = {} if iters is None else dict(nb_iters=it, b_iters=it)
            wv, wf = orc.marching_cubes(tagged, target, mn, mx, **kw)
            counts = np.zeros(2, np.uint32)
            gv, gf = np.zeros((len(wv) + 8, 4), np.float32), np.zeros((len(wf) + 8, 4), np.uint32)
            assert glsl.glsl_marching_cubes(padded, d, target, mn, mx, it, 0.9, it, 0.2, gv.ctypes.data, len(gv), gf.ctypes.data, len(gf), counts) == 0
            assert (int(counts[0]), int(counts[1])) == (len(wv), len(wf)) and len(wv) > 0
            assert np.array_equal(gf[: len(wf)], wf), f"faces, target {target}"
            assert np.array_equal(gv[: len(wv)].view(np.uint32), wv.view(np.uint32)), f"vertices, target {target}"
