"""Synthetic inputs for tests and benchmarks (input synthesis only — no part of the hot path is computed here).

The reference ships no mesh generator and no mesh fixture (its samples were stripped, .MISSING_LARGE_BLOBS); SURVEY.md §8(d)
defines the stand-in used for every config: a closed, double-walled vessel of revolution about y with ~20k triangles,
normalised exactly like CADModel::load (SRC/Graphics/Core/CADModel.cpp:148-152: scale 0.499999*2/max(extent), centred)."""
from __future__ import annotations

import numpy as np

MODEL_NORMALIZATION_SCALE = np.float32(0.499999)  # CADModel.cpp:20


def _uniform_stream(seed: int):
    """std::mt19937(seed) + the libstdc++ float recipe (SURVEY finding 9) on numpy's identical MT19937 core."""
    rs = np.random.RandomState(seed)

    def draw() -> np.float32:
        u = np.float32(int(rs.randint(0, 2**32, dtype=np.uint64))) * np.float32(2.3283064365386963e-10)
        return np.float32(0.99999994) if u >= np.float32(1.0) else u

    return draw


def vessel_mesh(mesh_idx: int = 0, n_ang: int = 100, n_prof: int = 50):
    """-> (vertices float32[nv][3], faces uint32[nf][3]); 19 800 triangles at the default resolution."""
    u = _uniform_stream(1000 + mesh_idx)
    base = np.float32(0.18) * (1 + np.float32(0.2) * (u() - np.float32(0.5)))
    a1 = np.float32(0.14) * (1 + np.float32(0.3) * (u() - np.float32(0.5)))
    a2 = np.float32(0.06) * (1 + np.float32(0.6) * (u() - np.float32(0.5)))
    height = np.float32(0.9) * (1 + np.float32(0.2) * (u() - np.float32(0.5)))
    wall = np.float32(0.035)
    t = np.linspace(0.0, 1.0, n_prof + 1, dtype=np.float64)
    r_o = (base + a1 * np.sin(np.pi * t) ** 0.8 + a2 * np.sin(3 * np.pi * t)).astype(np.float32)
    th = (np.arange(n_ang, dtype=np.float64) * (2 * np.pi / n_ang))
    cs, sn = np.cos(th).astype(np.float32), np.sin(th).astype(np.float32)
    j0 = int(np.ceil(0.06 * n_prof))  # inner wall starts at t = 0.06

    verts, faces = [], []

    def ring(r, y):
        i0 = len(verts)
        for i in range(n_ang):
            verts.append((r * cs[i], y, r * sn[i]))
        return i0

    outer = [ring(r_o[j], np.float32(height * np.float32(t[j]))) for j in range(n_prof + 1)]
    inner = {j: ring(r_o[j] - wall, np.float32(height * np.float32(t[j]))) for j in range(j0, n_prof + 1)}

    def band(a, b, flip):
        for i in range(n_ang):
            i1 = (i + 1) % n_ang
            q = (a + i, a + i1, b + i1, b + i)
            if flip:
                faces.append((q[0], q[2], q[1])), faces.append((q[0], q[3], q[2]))
            else:
                faces.append((q[0], q[1], q[2])), faces.append((q[0], q[2], q[3]))

    for j in range(n_prof):
        band(outer[j], outer[j + 1], False)
    for j in range(j0, n_prof):
        band(inner[j], inner[j + 1], True)
    band(outer[n_prof], inner[n_prof], False)  # annular rim
    c_out = len(verts)
    verts.append((np.float32(0), np.float32(0), np.float32(0)))  # outer bottom disc
    c_in = len(verts)
    verts.append((np.float32(0), np.float32(height * np.float32(t[j0])), np.float32(0)))  # inner floor
    for i in range(n_ang):
        i1 = (i + 1) % n_ang
        faces.append((c_out, outer[0] + i1, outer[0] + i))
        faces.append((c_in, inner[j0] + i, inner[j0] + i1))

    v = np.asarray(verts, dtype=np.float32)
    f = np.asarray(faces, dtype=np.uint32)
    # CADModel::load normalisation (CADModel.cpp:148-152), float32
    mn, mx = v.min(0), v.max(0)
    center = (mx + mn) / np.float32(2)
    extent = mx - center
    scale = MODEL_NORMALIZATION_SCALE / extent.max() * np.float32(2)
    v = ((v - center) * scale).astype(np.float32)
    return np.ascontiguousarray(v), np.ascontiguousarray(f)


def mesh_aabb(verts):
    return verts.min(0).astype(np.float32), verts.max(0).astype(np.float32)


def solid_vessel_params(mesh_idx: int = 0):
    """(base, a1, a2) of the analytic solid's outer profile, per-mesh perturbed like vessel_mesh."""
    u = _uniform_stream(1000 + mesh_idx)
    base = np.float32(0.18) * (1 + np.float32(0.2) * (u() - np.float32(0.5)))
    a1 = np.float32(0.14) * (1 + np.float32(0.3) * (u() - np.float32(0.5)))
    a2 = np.float32(0.06) * (1 + np.float32(0.6) * (u() - np.float32(0.5)))
    return float(base), float(a1), float(a2)


def solid_vessel_inside(x, y, z, params):
    """analytic inside-test at normalised coordinates (x, z in [-.5, .5], y in [0, 1]); numpy arrays or scalars"""
    base, a1, a2 = params
    r = np.sqrt(x * x + z * z) / 0.6
    ro = base + a1 * np.maximum(np.sin(np.pi * y), 0) ** 0.8 + a2 * np.sin(3 * np.pi * y)
    return (r <= ro) & ((r >= ro - 0.06) | (y < 0.08))


def solid_vessel_seeds(n: int, nseeds: int, params, seed: int = 80):
    """nseeds distinct cells inside the analytic solid of an n^3 grid, by rejection sampling with the reference RNG recipe
    (three draws per attempt), sorted like Seeder::uniform's std::set, labels 2.."""
    u = _uniform_stream(seed)
    got = set()
    while len(got) < nseeds:
        x, y, z = (int(np.float32(n - 1) * u()) for _ in range(3))
        if solid_vessel_inside((x + 0.5) / n - 0.5, (y + 0.5) / n, (z + 0.5) / n - 0.5, params):
            got.add((x, y, z))
    pts = sorted(got)
    return np.array([[x, y, z, 2 + i] for i, (x, y, z) in enumerate(pts)], dtype=np.uint32)


def solid_vessel_grid(n: int, mesh_idx: int = 0) -> np.ndarray:
    """Analytic inside-test of the same vessel wall (between outer and inner profile) sampled at cell centres of an n^3
    grid: the solid used for cfg5-style flood tests where a 20k-triangle SAT at billions of cells is pointless."""
    u = _uniform_stream(1000 + mesh_idx)
    base = np.float32(0.18) * (1 + np.float32(0.2) * (u() - np.float32(0.5)))
    a1 = np.float32(0.14) * (1 + np.float32(0.3) * (u() - np.float32(0.5)))
    a2 = np.float32(0.06) * (1 + np.float32(0.6) * (u() - np.float32(0.5)))
    c = (np.arange(n, dtype=np.float64) + 0.5) / n
    x, y, z = np.meshgrid(c - 0.5, c, c - 0.5, indexing="ij", sparse=True)
    r = np.sqrt(x * x + z * z) / 0.6
    ro = base + a1 * np.sin(np.pi * y) ** 0.8 + a2 * np.sin(3 * np.pi * y)
    inside = (r <= ro) & ((r >= ro - 0.06) | (y < 0.08))
    return inside.astype(np.uint16)
