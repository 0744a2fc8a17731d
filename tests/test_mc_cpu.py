"""f2 on the host: the packed marching-cubes case table (consistency everywhere, identity with the reference's copy where /root/reference
exists) and properties of the oracle's marching cubes + fusion + smoothing that do not depend on any other implementation."""
import os
import re
from collections import Counter

import numpy as np

from conftest import ROOT, random_blob_grid

EDGES = [(0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4), (0, 4), (1, 5), (2, 6), (3, 7)]  # marchingCubes-comp.glsl:40-54


def _rows():
    text = open(os.path.join(ROOT, "voxelfragmentml_b200", "csrc", "mc_tritable.inc")).read()
    words = [int(w, 16) for w in re.findall(r"0x([0-9A-F]{16})ull", text)]
    assert len(words) == 256
    rows = []
    for w in words:
        nib = [(w >> (4 * k)) & 0xF for k in range(16)]
        n = nib.index(0xF)
        assert n % 3 == 0 and n <= 15 and all(v == 0xF for v in nib[n:])
        rows.append(nib[:n])
    return rows


def test_case_table_is_consistent_with_the_corner_numbering():
    rows = _rows()
    for case, row in enumerate(rows):
        cut = {e for e, (a, b) in enumerate(EDGES) if ((case >> a) & 1) != ((case >> b) & 1)}
        assert set(row) <= cut and all(v < 12 for v in row), case           # a triangle only uses edges its case cuts
        assert (len(row) == 0) == (case in (0, 255))
        assert set(row) == cut                                               # ... and every cut edge carries surface
    assert max(len(r) for r in rows) == 15


def test_case_table_equals_the_reference_copy():
    src = "/root/reference/MeshFragments/Source/Graphics/Core/MarchingCubes.cpp"
    if not os.path.exists(src):
        import pytest

        pytest.skip("reference tree not present")
    text = open(src).read()
    body = text[text.index("_triangleTable[256 * 16]"):]
    vals = [int(v) for v in re.findall(r"-?\d+", body[body.index("{") + 1:body.index("};")])]
    want = [[v for v in vals[16 * i:16 * i + 16] if v >= 0] for i in range(256)]
    assert _rows() == want


def _edge_use(f):
    c = Counter()
    for t in f[:, :3]:
        for a, b in ((t[0], t[1]), (t[1], t[2]), (t[2], t[0])):
            c[(min(int(a), int(b)), max(int(a), int(b)))] += 1
    return c


def test_fragment_surface_is_closed_and_fused(orc):
    """every fragment of a labelled blob gives a closed 2-manifold-ish surface: each edge is shared by an even number of faces, no
    duplicated vertex positions before smoothing, flags in {0, 1}, coordinates inside the AABB grown by one cell."""
    occ = random_blob_grid((28, 24, 30), 3)
    seeds = np.uint32([[5, 5, 5, 2], [20, 18, 22, 3], [12, 20, 8, 4]])
    lab = orc.detect_boundaries(orc.naive((occ != 0).astype(np.uint16), seeds, 0), 1)
    mn, mx = np.float32([-0.4, -0.3, -0.5]), np.float32([0.4, 0.3, 0.5])
    for target in (2, 3, 4):
        v, f = orc.marching_cubes(lab, target, mn, mx, nb_iters=0, b_iters=0)
        assert len(f) > 0 and f[:, :3].max() == len(v) - 1
        assert all(n % 2 == 0 for n in _edge_use(f).values())
        assert len(np.unique(v[:, :3], axis=0)) == len(v)
        assert set(np.unique(v[:, 3])) <= {0.0, 1.0} and set(np.unique(f[:, 3])) <= {0, 1}
        cell = (mx - mn) / np.float32(lab.shape)
        assert (v[:, :3] >= mn - cell - 1e-6).all() and (v[:, :3] <= mx + cell + 1e-6).all()
        # a face is a boundary face iff one of its vertices is
        assert np.array_equal(f[:, 3], v[f[:, :3], 3].max(axis=1).astype(np.uint32))
        vs, fs = orc.marching_cubes(lab, target, mn, mx, nb_iters=4, b_iters=4)
        assert np.array_equal(fs, f) and not np.array_equal(vs[:, :3], v[:, :3])  # smoothing moves vertices, never the connectivity


def test_single_voxel_and_absent_label(orc):
    g = np.zeros((5, 5, 5), np.uint16)
    g[2, 2, 2] = 7
    v, f = orc.marching_cubes(g, 7, np.float32([0, 0, 0]), np.float32([5, 5, 5]), nb_iters=0, b_iters=0)
    assert len(v) == 6 and len(f) == 8  # an octahedron around the cell
    # padded-grid vertex p maps to p * 1 + (0 - 1): the six edge midpoints around cell (2, 2, 2) whose padded centre is 3
    assert sorted(map(tuple, v[:, :3].tolist())) == sorted([(1.5, 2.0, 2.0), (2.5, 2.0, 2.0), (2.0, 1.5, 2.0), (2.0, 2.5, 2.0), (2.0, 2.0, 1.5), (2.0, 2.0, 2.5)])
    v, f = orc.marching_cubes(g, 9, np.float32([0, 0, 0]), np.float32([5, 5, 5]))
    assert len(v) == 0 and len(f) == 0


def test_fragment_touching_the_grid_faces_is_closed_by_the_padding(orc):
    g = np.full((6, 4, 5), 2, np.uint16)
    v, f = orc.marching_cubes(g, 2, np.float32([0, 0, 0]), np.float32([6, 4, 5]), nb_iters=0, b_iters=0)
    assert all(n == 2 for n in _edge_use(f).values())
    assert len(v) - len(_edge_use(f)) + len(f) == 2  # Euler characteristic of a sphere


# ---- the marching stage restated a second time, literally from the shader text (marchingCubes-comp.glsl:57-168 + MarchingCubes::setGrid,
# MarchingCubes.cpp:516-533 + the transformation of RegularGrid::toTriangleMesh, RegularGrid.cpp:477-479), compared with the oracle as an
# ORDER-FREE multiset of oriented triangles: the reference's vertex / face order is an atomicAdd race, geometry is not.
NEIGHBORS = [(0, 0, 0), (0, 0, 1), (-1, 0, 1), (-1, 0, 0), (0, 1, 0), (0, 1, 1), (-1, 1, 1), (-1, 1, 0)]  # marchingCubes-comp.glsl:27-37


def _march_literal(grid, target):
    """triangles in PADDED-grid coordinates doubled (all coordinates are half-integers: the field is 0/1 and isolevel 0.5, so mu = 0.5)"""
    X, Y, Z = (s + 2 for s in grid.shape)
    pad = np.full((X, Y, Z), 1, np.uint16)  # setGrid: border = VOXEL_FREE
    pad[1:-1, 1:-1, 1:-1] = grid
    field = ((pad & 0x7FFF) == target).astype(np.float32)  # :117 unmaskedBit(value, 15) == targetValue
    rows = _rows()
    tris = []
    for x in range(X):
        for y in range(Y):
            for z in range(Z):
                if x == 0 or y == Y - 1 or z == Z - 1:  # :106-107
                    continue
                values = [field[x + n[0], y + n[1], z + n[2]] for n in NEIGHBORS]
                configuration = sum(1 << i for i in range(8) if values[i] < 0.5)  # :119-120
                if not rows[configuration]:
                    continue
                vlist = {}
                for i, (a, b) in enumerate(EDGES):
                    if (values[a] < 0.5) != (values[b] < 0.5):  # configurationTable[configuration] & (1 << i): the cut edges
                        p1 = np.float32([x + NEIGHBORS[a][0], y + NEIGHBORS[a][1], z + NEIGHBORS[a][2]])
                        p2 = np.float32([x + NEIGHBORS[b][0], y + NEIGHBORS[b][1], z + NEIGHBORS[b][2]])
                        mu = (np.float32(0.5) - values[a]) / (values[b] - values[a])  # findVertex :63-79 (neither EPSILON branch can fire)
                        vlist[i] = p1 + mu * (p2 - p1)
                row = rows[configuration]
                for i in range(0, len(row), 3):  # :140-151: vertices 0, 2, 1 of the table row
                    tri = [vlist[row[i]], vlist[row[i + 2]], vlist[row[i + 1]]]
                    tris.append(tuple(tuple(int(round(2 * float(c))) for c in p) for p in tri))
    return tris


def _canon(tri):  # oriented triangle up to rotation of its vertices
    k = tri.index(min(tri))
    return tri[k:] + tri[:k]


def test_marching_stage_matches_the_shader_text(orc):
    occ = random_blob_grid((13, 11, 14), 5)
    seeds = np.uint32([[3, 3, 3, 2], [9, 8, 10, 3], [6, 9, 4, 4]])
    lab = orc.detect_boundaries(orc.naive((occ != 0).astype(np.uint16), seeds, 0), 1)  # tagged words: the shader unmasks bit 15
    lab[0, :, :] = np.where(lab[0, :, :] > 1, lab[0, :, :], 2)                      # a fragment that touches a grid face
    mn, mx = np.float32([-0.45, -0.25, -0.5]), np.float32([0.35, 0.3, 0.45])
    scale = (mx - mn) / np.float32(lab.shape)
    t = mn - scale  # translate(-scale) * translate(minPoint) * scale(scale)
    for target in (2, 3, 4):
        v, f = orc.marching_cubes(lab, target, mn, mx, nb_iters=0, b_iters=0)
        q = np.rint(2.0 * (v[:, :3].astype(np.float64) - t) / scale).astype(np.int64)  # back to doubled padded-grid coordinates
        assert np.abs(2.0 * (v[:, :3].astype(np.float64) - t) / scale - q).max() < 1e-3
        got = Counter(_canon(tuple(tuple(int(c) for c in q[i]) for i in tri[:3])) for tri in f)
        want = Counter(_canon(tri) for tri in _march_literal(lab, target))
        assert sum(want.values()) > 50 and got == want, target


# ---- the smoothing passes transcribed from resetLaplacianBuffer / laplacianSmoothing / finishLaplacianSmoothing-comp.glsl and the host loop
# MarchingCubes::smoothSurface (MarchingCubes.cpp:498-521; called non-boundary first, then boundary, :399-400).  Integer atomics: the result does
# not depend on invocation order, so the oracle must reproduce it bit for bit given the same fused mesh.
def _smooth_literal(v, f, iters, weight, boundary):
    f32 = np.float32
    v = v.copy()
    EPS, UINT_MULT = f32(0.00000001), f32(10000.0)
    target = f32(1.0 if boundary else 0.0)
    check_validity = 0 if boundary else 1
    for _ in range(iters):
        lap = np.zeros((len(v), 4), np.int64)
        for face in f:
            valid = True
            if check_validity == 1:
                for i in range(3):
                    valid = valid and abs(v[face[i], 3] - target) < EPS
            if not valid:
                continue
            for i in range(3):
                vertex = [int(c) for c in (v[face[i], :3] * UINT_MULT)] + [1]  # ivec3(xyz * UINT_MULT): truncation towards zero
                for other in (face[(i + 1) % 3], face[(i + 2) % 3]):
                    lap[other] += vertex
        lap = ((lap + 2 ** 31) % 2 ** 32 - 2 ** 31).astype(np.int32)  # the buffers are 32-bit ints
        for idx in range(len(v)):
            if abs(v[idx, 3] - target) < EPS and lap[idx, 3] > 0:
                mean = [f32(lap[idx, k]) / f32(lap[idx, 3]) / UINT_MULT for k in range(3)]
                for k in range(3):  # mix(a, b, w) = a * (1 - w) + b * w
                    v[idx, k] = f32(v[idx, k] * f32(f32(1.0) - f32(weight))) + f32(mean[k] * f32(weight))
    return v


def test_smoothing_matches_the_shader_text(orc):
    occ = random_blob_grid((12, 10, 13), 7)
    seeds = np.uint32([[3, 3, 3, 2], [8, 7, 9, 3]])
    lab = orc.detect_boundaries(orc.naive((occ != 0).astype(np.uint16), seeds, 0), 1)
    mn, mx = np.float32([-0.45, -0.25, -0.5]), np.float32([0.35, 0.3, 0.45])
    for target in (2, 3):
        v0, f = orc.marching_cubes(lab, target, mn, mx, nb_iters=0, b_iters=0)
        assert 0 < v0[:, 3].sum() < len(v0)  # both vertex kinds occur
        for nb, wnb, b, wb in ((2, 0.9, 0, 0.2), (0, 0.9, 3, 0.2), (2, 0.9, 2, 0.2)):
            vs, fs = orc.marching_cubes(lab, target, mn, mx, nb_iters=nb, nb_weight=wnb, b_iters=b, b_weight=wb)
            want = _smooth_literal(_smooth_literal(v0, f[:, :3], nb, wnb, False), f[:, :3], b, wb, True)
            assert np.array_equal(fs, f)
            assert np.array_equal(vs.view(np.uint32), want.view(np.uint32)), (target, nb, b)
