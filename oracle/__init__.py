"""ctypes loader for the CPU ORACLE (oracle/vf_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs.  The product package (voxelfragmentml_b200/) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libvf_oracle.so")

EUCLIDEAN, MANHATTAN, CHEBYSHEV = 0, 1, 2
INNER, OUTER, BOTH = 0, 1, 2
STD_UNIFORM, HALTON, BOOST_NORMAL = 0, 1, 2
SQUARE, ELLIPSE, CROSS = 0, 1, 2
KEY_WALL, KEY_UNREACHED, KEY_DIST_SHIFT = 0xFFFFFFFF, 0xFFFFFFFE, 15


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (g++ only; no reference sources are copied)."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(
        os.path.getmtime(os.path.join(_HERE, f)) for f in ("vf_oracle.cpp", "vf_oracle.h", "Makefile")
    ):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _declare(_lib)
    return _lib


_u16p = np.ctypeslib.ndpointer(np.uint16, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")


class FloodStats(C.Structure):
    _fields_ = [("levels", C.c_uint32), ("rounds", C.c_uint32), ("freed_voxels", C.c_uint32), ("max_dist", C.c_uint32)]


def _declare(L):
    L.orc_rng_create.restype = C.c_void_p
    L.orc_rng_create.argtypes = [C.c_uint32]
    L.orc_rng_destroy.argtypes = [C.c_void_p]
    L.orc_rng_seed.argtypes = [C.c_void_p, C.c_uint32]
    L.orc_rng_raw.restype = C.c_uint32
    L.orc_rng_raw.argtypes = [C.c_void_p]
    L.orc_rng_uniform.restype = C.c_float
    L.orc_rng_uniform.argtypes = [C.c_void_p]
    L.orc_rng_uniform_range.restype = C.c_float
    L.orc_rng_uniform_range.argtypes = [C.c_void_p, C.c_float, C.c_float]
    L.orc_rng_uniform_int.restype = C.c_int
    L.orc_rng_uniform_int.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.orc_selfcheck_rng.restype = C.c_int
    L.orc_selfcheck_rng.argtypes = [C.c_uint32, C.c_int]
    L.orc_decode_position.argtypes = [C.c_uint32, _u32p, C.c_int, _u32p]
    L.orc_dims_rule.argtypes = [_f32p, _f32p, C.c_uint32, _u32p]
    L.orc_tri_box_intersect.restype = C.c_int
    L.orc_tri_box_intersect.argtypes = [_f32p] * 5
    L.orc_voxelize_sat.restype = C.c_int
    L.orc_voxelize_sat.argtypes = [_f32p, C.c_uint32, _u32p, C.c_uint32, _f32p, _f32p, _u32p, _u16p, C.c_int, C.c_void_p]
    L.orc_voxelize_solid.restype = C.c_int
    L.orc_voxelize_solid.argtypes = [_f32p, C.c_uint32, _u32p, C.c_uint32, _f32p, _f32p, _u32p, _u16p, C.c_int]
    L.orc_seed_uniform.restype = C.c_int
    L.orc_seed_uniform.argtypes = [C.c_void_p, _u16p, _u32p, C.c_uint32, C.c_int, C.c_int, _u32p, C.POINTER(C.c_uint32)]
    L.orc_merge_seeds.argtypes = [_u32p, C.c_uint32, _u32p, C.c_uint32, C.c_int]
    L.orc_make_seeds.restype = C.c_int
    L.orc_make_seeds.argtypes = [C.c_void_p, _u16p, _u32p, C.c_uint32, C.c_uint32, C.c_int, C.c_int, _u32p, C.c_uint32]
    L.orc_near_seeds.restype = C.c_int
    L.orc_near_seeds.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_uint32), _u16p, _u32p, _u32p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                 _u32p, C.c_uint32]
    L.orc_naive.argtypes = [_u16p, _u32p, _u32p, C.c_uint32, C.c_int, C.c_int]
    L.orc_flood.restype = C.c_int
    L.orc_flood.argtypes = [_u16p, _u32p, _u32p, C.c_uint32, C.c_int, C.c_int, C.c_int, C.POINTER(FloodStats)]
    L.orc_flood_keys.restype = C.c_int
    L.orc_flood_keys.argtypes = [_u16p, _u32p, _u32p, C.c_uint32, C.c_int, _u32p]
    L.orc_relax_keys_slab.restype = C.c_uint64
    L.orc_relax_keys_slab.argtypes = [_u32p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int]
    L.orc_remove_isolated_regions_cpu.argtypes = [_u16p, _u32p, _u32p, C.c_uint32]
    L.orc_detect_boundaries.argtypes = [_u16p, _u32p, C.c_int]
    L.orc_fill_noise.argtypes = [C.c_void_p, _f32p, C.c_uint32]
    L.orc_erode.argtypes = [_u16p, _u32p, C.c_int, C.c_uint32, C.c_uint32, C.c_float, C.c_float, _f32p, C.c_uint32, C.c_int]
    L.orc_erode_mask.argtypes = [C.c_int, C.c_uint32, _f32p, C.POINTER(C.c_float), C.POINTER(C.c_uint32)]
    L.orc_remove_isolated_regions_grid.argtypes = [_u16p, _u32p]
    L.orc_undo_mask.argtypes = [_u16p, C.c_uint64, C.c_uint32, C.c_int]
    L.orc_reset_filling.argtypes = [_u16p, C.c_uint64]
    L.orc_homogenize.argtypes = [_u16p, C.c_uint64]
    L.orc_count_values.restype = C.c_uint64
    L.orc_count_values.argtypes = [_u16p, C.c_uint64, _u32p]
    L.orc_encode_rle.restype = C.c_uint64
    L.orc_encode_rle.argtypes = [_u16p, _u32p, C.c_void_p, C.c_uint64]
    L.orc_decode_rle.restype = C.c_int
    L.orc_decode_rle.argtypes = [C.c_char_p, C.c_uint64, _u32p, C.c_void_p, C.c_uint64]
    L.orc_halton.restype = C.c_float
    L.orc_halton.argtypes = [C.c_uint, C.c_uint]
    L.orc_encode_vox.restype = C.c_uint64
    L.orc_encode_vox.argtypes = [_u16p, _u32p, C.c_int, C.c_void_p, C.c_uint64]
    L.orc_encode_qstack.restype = C.c_uint64
    L.orc_encode_qstack.argtypes = [_u16p, _u32p, C.c_void_p, C.c_uint64]
    L.orc_encode_bing_squared.restype = C.c_uint64
    L.orc_encode_bing_squared.argtypes = [_u16p, _u32p, C.c_void_p, C.c_uint64]
    L.orc_mc_soup.restype = C.c_uint32
    L.orc_mc_soup.argtypes = [_u16p, _u32p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32]
    L.orc_marching_cubes.restype = C.c_int
    L.orc_marching_cubes.argtypes = [_u16p, _u32p, C.c_uint32, _f32p, _f32p, C.c_uint32, C.c_float, C.c_uint32, C.c_float, C.c_void_p, C.c_uint32,
                                     C.c_void_p, C.c_uint32, _u32p]
    L.orc_num_threads.restype = C.c_int
    L.orc_set_num_threads.argtypes = [C.c_int]


class OracleError(RuntimeError):
    def __init__(self, code, what):
        super().__init__(f"oracle: {what} failed with code {code}")
        self.code = code


def _dims(grid_or_dims) -> np.ndarray:
    if isinstance(grid_or_dims, np.ndarray) and grid_or_dims.ndim == 3:
        return np.asarray(grid_or_dims.shape, dtype=np.uint32)
    return np.ascontiguousarray(grid_or_dims, dtype=np.uint32)


def _seeds(seeds) -> np.ndarray:
    s = np.ascontiguousarray(seeds, dtype=np.uint32)
    assert s.ndim == 2 and s.shape[1] == 4
    return s


class Rng:
    """std::mt19937 + the libstdc++ float recipe (SURVEY finding 9)."""

    def __init__(self, seed: int = 80):
        self._h = lib().orc_rng_create(seed)

    def __del__(self):
        try:
            lib().orc_rng_destroy(self._h)
        except Exception:
            pass

    def seed(self, s):
        lib().orc_rng_seed(self._h, s)

    def raw(self):
        return lib().orc_rng_raw(self._h)

    def uniform(self):
        return lib().orc_rng_uniform(self._h)

    def uniform_int(self, lo, hi):
        return lib().orc_rng_uniform_int(self._h, lo, hi)

    def fill_noise(self, n: int) -> np.ndarray:
        out = np.empty(n, dtype=np.float32)
        lib().orc_fill_noise(self._h, out, n)
        return out


def selfcheck_rng(seed=80, ndraws=100000) -> int:
    return lib().orc_selfcheck_rng(seed, ndraws)


def decode_position(index, dims, mode=0):
    out = np.zeros(3, dtype=np.uint32)
    lib().orc_decode_position(index, _dims(dims), mode, out)
    return tuple(int(v) for v in out)


def dims_rule(aabb_min, aabb_max, max_voxels):
    out = np.zeros(3, dtype=np.uint32)
    lib().orc_dims_rule(np.ascontiguousarray(aabb_min, np.float32), np.ascontiguousarray(aabb_max, np.float32), max_voxels, out)
    return tuple(int(v) for v in out)


def tri_box_intersect(p1, p2, p3, bmin, bmax) -> bool:
    a = [np.ascontiguousarray(v, np.float32) for v in (p1, p2, p3, bmin, bmax)]
    return bool(lib().orc_tri_box_intersect(*a))


def voxelize_sat(verts, faces, aabb_min, aabb_max, dims, want_margin=False):
    verts = np.ascontiguousarray(verts, np.float32)
    faces = np.ascontiguousarray(faces, np.uint32)
    d = _dims(dims)
    grid = np.zeros(tuple(int(v) for v in d), dtype=np.uint16)
    margin = np.empty(grid.shape, dtype=np.float32) if want_margin else None
    rc = lib().orc_voxelize_sat(verts, len(verts), faces, len(faces), np.ascontiguousarray(aabb_min, np.float32),
                                np.ascontiguousarray(aabb_max, np.float32), d, grid, 1,
                                margin.ctypes.data if want_margin else None)
    if rc:
        raise OracleError(rc, "voxelize_sat")
    return (grid, margin) if want_margin else grid


def voxelize_solid(verts, faces, aabb_min, aabb_max, dims):
    """Tetravoxelizer occupancy (the reference's live RegularGrid::fill), rasteriser rule as fixed in vf_oracle.cpp"""
    verts = np.ascontiguousarray(verts, np.float32)
    faces = np.ascontiguousarray(faces, np.uint32)
    d = _dims(dims)
    grid = np.zeros(tuple(int(v) for v in d), dtype=np.uint16)
    rc = lib().orc_voxelize_solid(verts, len(verts), faces, len(faces), np.ascontiguousarray(aabb_min, np.float32),
                                  np.ascontiguousarray(aabb_max, np.float32), d, grid, 1)
    if rc:
        raise OracleError(rc, "voxelize_solid")
    return grid


def seed_uniform(rng: Rng, grid, n, mode=STD_UNIFORM, location=OUTER):
    out = np.zeros((n, 4), dtype=np.uint32)
    att = C.c_uint32(0)
    rc = lib().orc_seed_uniform(rng._h, grid, _dims(grid), n, mode, location, out, C.byref(att))
    if rc:
        raise OracleError(rc, "seed_uniform")
    return out, att.value


def merge_seeds(frags, seeds, dfunc=EUCLIDEAN):
    s = _seeds(seeds).copy()
    f = _seeds(frags)
    lib().orc_merge_seeds(f, len(f), s, len(s), dfunc)
    return s


def make_seeds(rng: Rng, grid, n, n_extra=0, mode=STD_UNIFORM, merge_dfunc=EUCLIDEAN):
    cap = n + (n + n_extra if n_extra else 0)
    out = np.zeros((cap, 4), dtype=np.uint32)
    rc = lib().orc_make_seeds(rng._h, grid, _dims(grid), n, n_extra, mode, merge_dfunc, out, cap)
    if rc < 0:
        raise OracleError(rc, "make_seeds")
    return out[:rc]


def near_seeds(rng: Rng, grid, frags, num_impacts, num_seeds, spreading, crand_state=80, crand_mode=0):
    """Seeder::nearSeeds; returns (seeds, crand_state afterwards).  crand_mode 0 = MSVC LCG (the reference's platform), 1 = libc rand()."""
    frags = _seeds(frags)
    out = np.zeros((len(frags) + num_seeds, 4), dtype=np.uint32)
    st = C.c_uint32(crand_state)
    rc = lib().orc_near_seeds(rng._h, crand_mode, C.byref(st), grid, _dims(grid), frags, len(frags), num_impacts, num_seeds, spreading, out, len(out))
    if rc < 0:
        raise OracleError(rc, "near_seeds")
    return out[:rc], int(st.value)


def naive(grid, seeds, dfunc=EUCLIDEAN, decode_mode=0):
    s = _seeds(seeds)
    lib().orc_naive(grid, _dims(grid), s, len(s), dfunc, decode_mode)
    return grid


def flood(grid, seeds, dfunc=MANHATTAN, id_bits=8, algo=0):
    s = _seeds(seeds)
    st = FloodStats()
    rc = lib().orc_flood(grid, _dims(grid), s, len(s), dfunc, id_bits, algo, C.byref(st))
    if rc:
        raise OracleError(rc, "flood")
    return grid, st


def flood_keys(grid, seeds, dfunc=MANHATTAN):
    s = _seeds(seeds)
    keys = np.empty(grid.shape, dtype=np.uint32)
    rc = lib().orc_flood_keys(grid, _dims(grid), s, len(s), dfunc, keys)
    if rc:
        raise OracleError(rc, "flood_keys")
    return keys


def relax_keys_slab(keys, nneigh):
    assert keys.ndim == 3 and keys.dtype == np.uint32
    return int(lib().orc_relax_keys_slab(keys, keys.shape[0], keys.shape[1], keys.shape[2], nneigh))


def remove_isolated_regions_cpu(grid, seeds):
    s = _seeds(seeds)
    lib().orc_remove_isolated_regions_cpu(grid, _dims(grid), s, len(s))
    return grid


def detect_boundaries(grid, boundary_size=1):
    lib().orc_detect_boundaries(grid, _dims(grid), boundary_size)
    return grid


def erode(grid, noise, type=ELLIPSE, size=3, iters=3, prob=0.5, thr=0.5, boundary_mode=0):
    noise = np.ascontiguousarray(noise, np.float32)
    lib().orc_erode(grid, _dims(grid), type, size, iters, prob, thr, noise, len(noise), boundary_mode)
    return grid


def erode_mask(type=ELLIPSE, size=3):
    k = size + (1 - size % 2)
    mask = np.zeros(k * k * k, dtype=np.float32)
    act = C.c_float(0)
    kk = C.c_uint32(0)
    lib().orc_erode_mask(type, size, mask, C.byref(act), C.byref(kk))
    return mask.reshape(k, k, k), act.value


def remove_isolated_regions_grid(grid):
    lib().orc_remove_isolated_regions_grid(grid, _dims(grid))
    return grid


def undo_mask(grid, position=15, rightmost=False):
    lib().orc_undo_mask(grid.reshape(-1), grid.size, position, int(rightmost))
    return grid


def reset_filling(grid):
    lib().orc_reset_filling(grid.reshape(-1), grid.size)
    return grid


def homogenize(grid):
    lib().orc_homogenize(grid.reshape(-1), grid.size)
    return grid


def count_values(grid):
    counts = np.zeros(32768, dtype=np.uint32)
    occ = lib().orc_count_values(grid.reshape(-1), grid.size, counts)
    return counts, int(occ)


def encode_rle(grid) -> bytes:
    d = _dims(grid)
    need = lib().orc_encode_rle(grid, d, None, 0)
    buf = np.empty(need, dtype=np.uint8)
    lib().orc_encode_rle(grid, d, buf.ctypes.data, need)
    return buf.tobytes()


def decode_rle(data: bytes) -> np.ndarray:
    d = np.frombuffer(data[:12], dtype=np.uint32).copy()
    grid = np.empty(tuple(int(v) for v in d), dtype=np.uint16)
    dd = np.zeros(3, dtype=np.uint32)
    rc = lib().orc_decode_rle(data, len(data), dd, grid.ctypes.data, grid.size)
    if rc:
        raise OracleError(rc, "decode_rle")
    return grid


def encode_bing_squared(grid) -> bytes:
    d = _dims(grid)
    need = lib().orc_encode_bing_squared(grid, d, None, 0)
    buf = np.empty(need, dtype=np.uint8)
    lib().orc_encode_bing_squared(grid, d, buf.ctypes.data, need)
    return buf.tobytes()


def halton(dimension: int, index: int) -> float:
    return float(lib().orc_halton(dimension, index))


def encode_vox(grid, squared: bool) -> bytes:
    d = _dims(grid)
    need = lib().orc_encode_vox(grid, d, int(bool(squared)), None, 0)
    buf = np.empty(need, dtype=np.uint8)
    lib().orc_encode_vox(grid, d, int(bool(squared)), buf.ctypes.data, need)
    return buf.tobytes()


def encode_qstack(grid) -> bytes:
    d = _dims(grid)
    need = lib().orc_encode_qstack(grid, d, None, 0)
    buf = np.empty(need, dtype=np.uint8)
    lib().orc_encode_qstack(grid, d, buf.ctypes.data, need)
    return buf.tobytes()


def marching_cubes(grid, target, aabb_min, aabb_max, nb_iters=None, nb_weight=0.9, b_iters=None, b_weight=0.2):
    """per-fragment marching cubes + vertex fusion + two-pass Laplacian smoothing; iterations default to unsigned(max dim * 0.048f)
    (FractureParameters.h:93,109; MarchingCubes.cpp:399-400).  Returns (vertices float32[nv][4], faces uint32[nf][4])."""
    d = _dims(grid)
    it = int(np.float32(max(int(v) for v in d)) * np.float32(0.048))
    nb_iters = it if nb_iters is None else nb_iters
    b_iters = it if b_iters is None else b_iters
    mn, mx = np.ascontiguousarray(aabb_min, np.float32), np.ascontiguousarray(aabb_max, np.float32)
    counts = np.zeros(2, np.uint32)
    lib().orc_marching_cubes(grid, d, int(target), mn, mx, nb_iters, nb_weight, b_iters, b_weight, None, 0, None, 0, counts)
    v = np.zeros((int(counts[0]), 4), np.float32)
    f = np.zeros((int(counts[1]), 4), np.uint32)
    if len(v) and len(f):
        lib().orc_marching_cubes(grid, d, int(target), mn, mx, nb_iters, nb_weight, b_iters, b_weight, v.ctypes.data, len(v), f.ctypes.data, len(f), counts)
    return v, f


def mc_soup(grid, target):
    """the triangle soup of one fragment before vertex fusion (marchingCubes-comp.glsl over the padded grid) and its Morton codes
    (computeMortonCodes-comp.glsl): (vertices float32[n][4] in padded-grid cells + boundary flag, codes uint32[n]), 3 vertices per triangle"""
    d = _dims(grid)
    n = int(lib().orc_mc_soup(grid, d, int(target), None, None, 0))
    v, m = np.zeros((n, 4), np.float32), np.zeros(n, np.uint32)
    if n:
        lib().orc_mc_soup(grid, d, int(target), v.ctypes.data, m.ctypes.data, n)
    return v, m


def num_threads() -> int:
    return lib().orc_num_threads()


def use_all_cores() -> int:
    """torchrun exports OMP_NUM_THREADS=1: ask OpenMP for every host core the process may run on."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    lib().orc_set_num_threads(n)
    return n
