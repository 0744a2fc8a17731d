"""Host-side mirror of the reference's operator surface for the hot path, over the C ABI.

Names, argument meaning and error behaviour follow the reference:
  RegularGrid              SRC/DataStructures/RegularGrid.h:15-367
  fracturer::Seeder        SRC/Fracturer/Seeder.h:54-76
  fracturer::Fracturer     SRC/Fracturer/Fracturer.h:22-51  (NaiveFracturer, FloodFracturer singletons)
  FractureParameters       SRC/Graphics/Core/FractureParameters.h:14-146
Every method is one C-ABI call; all arithmetic runs in libvoxfrag's CUDA kernels."""
from __future__ import annotations

import ctypes as C
import enum
import weakref

import numpy as np

from . import _capi
from ._capi import VfFloodStats, VfParams, check, ptr

VOXEL_EMPTY = 0  # RegularGrid.h:12
VOXEL_FREE = 1   # RegularGrid.h:13


class DistanceFunction(enum.IntEnum):  # Fracturer.h:10-14
    EUCLIDEAN = 0
    MANHATTAN = 1
    CHEBYSHEV = 2


class FractureAlgorithm(enum.IntEnum):  # FractureParameters.h:17
    NAIVE = 0
    FLOOD = 1
    VORONOI = 2


class RandomUniformType(enum.IntEnum):  # FractureParameters.h:23
    STD_UNIFORM = 0
    HALTON = 1
    BOOST_NORMAL_DISTRIBUTION = 2


class ErosionType(enum.IntEnum):  # FractureParameters.h:26
    SQUARE = 0
    ELLIPSE = 1
    CROSS = 2


class ExportGrid(enum.IntEnum):  # FractureParameters.h:35
    RLE = 0
    QUADSTACK = 1
    VOX = 2
    UNCOMPRESSED_BINARY = 3


class Location(enum.IntEnum):  # Seeder.h:24
    INNER = 0
    OUTER = 1
    BOTH = 2


class FractureParameters:
    """struct FractureParameters with the reference's defaults (FractureParameters.h:91-145); attribute names carry the
    reference's leading underscore (``_numSeeds``) and map onto ``struct vf_params``."""

    def __init__(self, **kw):
        object.__setattr__(self, "_c", VfParams())
        _capi.load().vf_params_default(C.byref(self._c))
        for k, v in kw.items():
            setattr(self, k, v)

    def __getattr__(self, name):
        f = name.lstrip("_")
        c = object.__getattribute__(self, "_c")
        if not hasattr(c, f):
            raise AttributeError(name)
        v = getattr(c, f)
        return tuple(v) if f == "voxelizationSize" else v

    def __setattr__(self, name, value):
        f = name.lstrip("_")
        if not hasattr(self._c, f):
            raise AttributeError(name)
        if f == "voxelizationSize":
            for i in range(3):
                self._c.voxelizationSize[i] = int(value[i])
        else:
            setattr(self._c, f, type(getattr(self._c, f))(value))


class Context:
    """One GPU + one CUDA stream + scratch memory (replaces the GL context and shader singletons)."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self._lib = _capi.load()
        h = C.c_void_p()
        if stream is None:
            check(self._lib.vf_ctx_create(device, C.byref(h)))
        else:
            check(self._lib.vf_ctx_create_on_stream(device, C.c_void_p(stream), C.byref(h)))
        self._h = h
        self.device = device
        self._grids = weakref.WeakSet()  # grids must be destroyed before their context (they borrow its device and stream)

    def close(self):
        if getattr(self, "_h", None):
            for g in list(self._grids):
                g.close()
            self._lib.vf_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reserve(self, dims):  # Fracturer::prepareSSBOs
        check(self._lib.vf_ctx_reserve(self._h, *[int(d) for d in dims]))

    def setBlockingSync(self, on=True):
        """host waits sleep instead of spinning (more contexts than host cores); 2: they poll and yield the core between polls"""
        check(self._lib.vf_ctx_set_blocking_sync(self._h, int(on)))

    def setFloodLevels(self, levels: int):
        """distance window of a flood round (0 = default 16: best latency; 8: best throughput with several jobs per GPU)"""
        check(self._lib.vf_ctx_set_flood_levels(self._h, int(levels)))

    def setFloodFront(self, max_front_cells: int):
        """flood phases start at cell granularity on one thread-block cluster and move to the tiles when more than this many (cell, key)
        pairs are pending (default 8192, 0 = tiles only: the setting for many contexts per GPU); same labels"""
        check(self._lib.vf_ctx_set_flood_front(self._h, int(max_front_cells)))

    def setFloodMode(self, ctas_per_sm: int):
        """flood phases: 1..4 = one cooperative launch per phase with that many CTAs per SM (default 4), 0 = one launch per round; same labels"""
        check(self._lib.vf_ctx_set_flood_mode(self._h, int(ctas_per_sm)))

    def setC1Mode(self, mode: int):
        """C1 (removeIsolatedRegions): 0 = descent certificate on grids of >= 2^26 cells with the union-find as fallback (default), 1 = union-find
        only, 2 = the certificate on grids of any size; same result"""
        check(self._lib.vf_ctx_set_c1_mode(self._h, int(mode)))

    def synchronize(self):
        check(self._lib.vf_ctx_synchronize(self._h))

    @property
    def stream(self) -> int:
        return self._lib.vf_ctx_stream(self._h) or 0

    @property
    def host_waits(self) -> int:
        """times a call of this context made the host wait for the stream so far"""
        return int(self._lib.vf_ctx_host_waits(self._h))

    @property
    def kernel_launches(self) -> int:
        return int(self._lib.vf_ctx_kernel_launches(self._h))

    def timer_start(self):
        check(self._lib.vf_ctx_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = C.c_float(0)
        check(self._lib.vf_ctx_timer_stop(self._h, C.byref(ms)))
        return ms.value

    # RandomUtilities (process-global in the reference; per context here)
    def initSeed(self, seed: int):
        check(self._lib.vf_rng_seed(self._h, int(seed) & 0xFFFFFFFF))

    def getUniformRandom(self) -> float:
        return self._lib.vf_rng_uniform(self._h)

    def rng_raw(self) -> int:
        return self._lib.vf_rng_raw(self._h)

    def fillNoiseBuffer(self, n: int) -> np.ndarray:  # RegularGrid::fillNoiseBuffer
        out = np.empty(n, dtype=np.float32)
        check(self._lib.vf_fill_noise(self._h, ptr(out), n))
        return out


def _seed_array(seeds) -> np.ndarray:
    s = np.ascontiguousarray(seeds, dtype=np.uint32)
    if s.ndim != 2 or s.shape[1] != 4:
        raise ValueError("seeds must be [n][4] = x, y, z, label")
    return s


class RegularGrid:
    """Device-resident uint16 label grid (z fastest).  ``dims`` = numDivs."""

    def __init__(self, ctx: Context, dims, device_ptr: int | None = None):
        self.ctx = ctx
        self._lib = ctx._lib
        h = C.c_void_p()
        X, Y, Z = (int(d) for d in dims)
        if device_ptr is None:
            check(self._lib.vf_grid_create(ctx._h, X, Y, Z, C.byref(h)))
        else:
            check(self._lib.vf_grid_wrap(ctx._h, C.c_void_p(device_ptr), X, Y, Z, C.byref(h)))
        self._h = h
        ctx._grids.add(self)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.vf_grid_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- geometry / storage
    def getNumSubdivisions(self):
        d = (C.c_uint32 * 3)()
        check(self._lib.vf_grid_dims(self._h, d))
        return tuple(d)

    @property
    def shape(self):
        return self.getNumSubdivisions()

    def setAABB(self, aabb_min, aabb_max, dims):
        mn = np.ascontiguousarray(aabb_min, np.float32)
        mx = np.ascontiguousarray(aabb_max, np.float32)
        check(self._lib.vf_grid_set_aabb(self._h, ptr(mn), ptr(mx), *[int(d) for d in dims]))

    def ssbo(self) -> int:
        return self._lib.vf_grid_device_ptr(self._h)

    def updateSSBO(self, host: np.ndarray):
        """host -> device (RegularGrid::updateSSBO)."""
        host = np.ascontiguousarray(host, dtype=np.uint16)
        if tuple(host.shape) != self.shape and host.size != int(np.prod(self.shape)):
            raise ValueError("host grid shape mismatch")
        check(self._lib.vf_grid_upload(self._h, ptr(host)))

    def updateGrid(self, out: np.ndarray | None = None) -> np.ndarray:
        """device -> host (RegularGrid::updateGrid)."""
        if out is None:
            out = np.empty(self.shape, dtype=np.uint16)
        check(self._lib.vf_grid_download(self._h, ptr(out)))
        return out

    data = updateGrid

    def upload_async(self, pinned_host):
        check(self._lib.vf_grid_upload_async(self._h, ptr(pinned_host)))

    def upload_bits(self, host_bits):
        """occupancy as one bit per cell (np.packbits(..., bitorder="little") of the linear grid != 0): set -> FREE, clear -> EMPTY; async on the stream"""
        check(self._lib.vf_grid_upload_bits(self._h, ptr(host_bits)))

    def download_async(self, pinned_host):
        check(self._lib.vf_grid_download_async(self._h, ptr(pinned_host)))

    def fillValue(self, value: int):
        check(self._lib.vf_grid_fill(self._h, int(value)))

    # ---- V2
    def fill(self, vertices, faces):
        """RegularGrid::fill(Model3D*): SAT surface voxelization of a triangle mesh (host float32[nv][3], uint32[nf][3])."""
        v = np.ascontiguousarray(vertices, np.float32)
        f = np.ascontiguousarray(faces, np.uint32)
        check(self._lib.vf_voxelize(self._h, ptr(v), len(v), ptr(f), len(f)))

    # ---- V1
    def fillSolid(self, vertices, faces) -> int:
        """RegularGrid::fill(Model3D*) as the reference runs it today (Tetravoxelizer, Tetravoxelizer.cpp:198-315): solid
        occupancy by XOR-ed tetrahedron slices.  Returns the number of FREE cells written."""
        v = np.ascontiguousarray(vertices, np.float32)
        f = np.ascontiguousarray(faces, np.uint32)
        occ = C.c_uint64(0)
        check(self._lib.vf_voxelize_solid(self._h, ptr(v), len(v), ptr(f), len(f), C.byref(occ)))
        return int(occ.value)

    # ---- f2
    def triangulateField(self, targetValue: int, boundaryMCIterations=0.048, boundaryMCWeight=0.2, nonBoundaryMCIterations=0.048, nonBoundaryMCWeight=0.9,
                         marchingCubesSubdivisions=1):
        """MarchingCubes::triangulateFieldGPU for one fragment label (what RegularGrid::toTriangleMesh runs per value, RegularGrid.cpp:482-483):
        -> (vertices float32[nv][4] = xyz + boundary flag, faces uint32[nf][4] = three vertex numbers + boundary flag).
        marchingCubesSubdivisions is carried only, as in the reference (FractureParameters.h:56 is read nowhere; RegularGrid.cpp:423 passes 1)."""
        mp = _capi.VfMcParams(boundaryMCIterations, boundaryMCWeight, nonBoundaryMCIterations, nonBoundaryMCWeight, int(marchingCubesSubdivisions))
        h = C.c_void_p()
        check(self._lib.vf_marching_cubes(self._h, int(targetValue), C.byref(mp), C.byref(h)))
        try:
            nv, nf = C.c_uint32(0), C.c_uint32(0)
            check(self._lib.vf_mesh_counts(h, C.byref(nv), C.byref(nf)))
            v = np.zeros((nv.value, 4), np.float32)
            f = np.zeros((nf.value, 4), np.uint32)
            check(self._lib.vf_mesh_download(h, ptr(v) if nv.value else None, ptr(f) if nf.value else None))
        finally:
            self._lib.vf_mesh_destroy(h)
        return v, f

    # ---- C1..C4
    def detectBoundaries(self, boundarySize: int = 1):
        check(self._lib.vf_detect_boundaries(self._h, int(boundarySize)))

    def erode(self, erosionType, convolutionSize, numIterations, erosionProbability, erosionThreshold, noise=None, boundaryMode=0):
        if noise is None:
            noise = self.ctx.fillNoiseBuffer(1000000)  # RegularGrid.cpp:126
        noise = np.ascontiguousarray(noise, np.float32)
        check(self._lib.vf_erode(self._h, int(erosionType), int(convolutionSize), int(numIterations), float(erosionProbability),
                                 float(erosionThreshold), ptr(noise), len(noise), int(boundaryMode)))

    def removeIsolatedRegions(self):
        check(self._lib.vf_remove_isolated_regions_grid(self._h))

    def undoMask(self):
        check(self._lib.vf_undo_mask(self._h))

    def resetFilling(self):
        check(self._lib.vf_reset_filling(self._h))

    def homogenize(self):
        check(self._lib.vf_homogenize(self._h))

    # ---- H1
    def countValues(self):
        """-> (counts[32768] indexed by label, numOccupiedVoxels)."""
        counts = np.zeros(_capi.HISTOGRAM_BINS, dtype=np.uint32)
        occ = C.c_uint64(0)
        check(self._lib.vf_histogram(self._h, ptr(counts), C.byref(occ)))
        return counts, int(occ.value)

    def countValuesUndoMask(self):
        """countValues followed by undoMask — the grid side of CADScene::prepareScene (CADScene.cpp:813-832) — in one pass."""
        counts = np.zeros(_capi.HISTOGRAM_BINS, dtype=np.uint32)
        occ = C.c_uint64(0)
        check(self._lib.vf_histogram_undo_mask(self._h, ptr(counts), C.byref(occ)))
        return counts, int(occ.value)

    def numOccupiedVoxels(self) -> int:
        return self.countValues()[1]

    # ---- X1
    def encodeRLE(self) -> bytes:
        """the `.rle` byte stream of exportRLE (RegularGrid.cpp:672-714), runs found on the device; only the stream is downloaded"""
        need = C.c_uint64(0)
        check(self._lib.vf_grid_encode_rle(self._h, None, 0, C.byref(need)))
        buf = np.empty(need.value, np.uint8)
        check(self._lib.vf_grid_encode_rle(self._h, ptr(buf), buf.size, C.byref(need)))
        return buf.tobytes()

    def encodeRLE_into(self, out) -> int:
        """the `.rle` stream written into a caller-owned host buffer (numpy uint8 array or pinned torch tensor) in ONE call: returns the stream's
        size; when it exceeds the buffer nothing is written and the size needed comes back negated"""
        need = C.c_uint64(0)
        cap = out.numel() * out.element_size() if hasattr(out, "numel") else out.nbytes
        check(self._lib.vf_grid_encode_rle(self._h, ptr(out), cap, C.byref(need)))
        return int(need.value) if need.value <= cap else -int(need.value)

    def exportGrid(self, filename: str, squared: bool, exportType):
        check(self._lib.vf_export(self._h, filename.encode(), int(exportType), int(bool(squared))))


class Seeder:
    """fracturer::Seeder (all static in the reference)."""

    INNER, OUTER, BOTH = Location.INNER, Location.OUTER, Location.BOTH

    @staticmethod
    def uniform(grid: RegularGrid, nseeds: int, randomSeedFunction=RandomUniformType.STD_UNIFORM, location=Location.OUTER,
                return_attempts: bool = False):
        out = np.zeros((nseeds, 4), dtype=np.uint32)
        att = C.c_uint32(0)
        check(grid._lib.vf_seed_uniform(grid._h, int(nseeds), int(randomSeedFunction), int(location), ptr(out), C.byref(att)))
        return (out, att.value) if return_attempts else out

    @staticmethod
    def mergeSeeds(frags, seeds, dfunc=DistanceFunction.EUCLIDEAN):
        f = _seed_array(frags)
        s = _seed_array(seeds).copy()
        check(_capi.load().vf_merge_seeds(ptr(f), len(f), ptr(s), len(s), int(dfunc)))
        return s

    @staticmethod
    def nearSeeds(grid: RegularGrid, frags, numImpacts: int, numSeeds: int, spreading: int):
        """Seeder::nearSeeds (Seeder.cpp:49-113): frags + impact-biased boundary seeds (C rand() = the MSVC LCG kept by the context)."""
        f = _seed_array(frags)
        out = np.zeros((len(f) + int(numSeeds), 4), dtype=np.uint32)
        cnt = C.c_uint32(0)
        check(grid._lib.vf_seed_near(grid._h, ptr(f), len(f), int(numImpacts), int(numSeeds), int(spreading), ptr(out), len(out), C.byref(cnt)))
        return out[: cnt.value]

    @staticmethod
    def make(grid: RegularGrid, numSeeds: int, numExtraSeeds: int = 0, randomSeedFunction=RandomUniformType.STD_UNIFORM,
             mergeDFunc=DistanceFunction.EUCLIDEAN):
        """Seed block of CADScene::fractureModel (CADScene.cpp:626-655)."""
        cap = numSeeds + (numSeeds + numExtraSeeds if numExtraSeeds else 0)
        out = np.zeros((cap, 4), dtype=np.uint32)
        cnt = C.c_uint32(0)
        check(grid._lib.vf_make_seeds(grid._h, numSeeds, numExtraSeeds, int(randomSeedFunction), int(mergeDFunc), ptr(out), cap,
                                      C.byref(cnt)))
        return out[: cnt.value]


class _Fracturer:
    def __init__(self):
        self._dfunc = DistanceFunction.EUCLIDEAN

    def setDistanceFunction(self, dfunc) -> bool:
        if int(dfunc) not in (0, 1, 2):
            return False
        self._dfunc = DistanceFunction(int(dfunc))
        return True

    def init(self, fractParameters=None):
        pass

    def prepareSSBOs(self, fractParameters=None, ctx: Context | None = None):
        if ctx is not None and fractParameters is not None:
            ctx.reserve(fractParameters._voxelizationSize)

    def destroy(self):
        pass


class NaiveFracturer(_Fracturer):
    """fracturer::NaiveFracturer (NaiveFracturer.cpp:215-225)."""

    def build(self, grid: RegularGrid, seeds, fractParameters: FractureParameters | None = None):
        s = _seed_array(seeds)
        check(grid._lib.vf_fracture_naive(grid._h, ptr(s), len(s), int(self._dfunc)))
        if fractParameters is not None and fractParameters._removeIsolatedRegions:
            check(grid._lib.vf_remove_isolated_regions(grid._h, ptr(s), len(s)))

    @staticmethod
    def removeIsolatedRegions(grid: RegularGrid, seeds):
        s = _seed_array(seeds)
        check(grid._lib.vf_remove_isolated_regions(grid._h, ptr(s), len(s)))


class FloodFracturer(_Fracturer):
    """fracturer::FloodFracturer (FloodFracturer.cpp:98-191)."""

    def __init__(self):
        super().__init__()
        self._dfunc = DistanceFunction.MANHATTAN  # FloodFracturer.cpp:29
        self.last_stats = None

    def build(self, grid: RegularGrid, seeds, fractParameters: FractureParameters | None = None, id_bits: int | None = None):
        s = _seed_array(seeds)
        if id_bits is None:
            id_bits = fractParameters._floodIdBits if fractParameters is not None else 0
        st = VfFloodStats()
        check(grid._lib.vf_fracture_flood(grid._h, ptr(s), len(s), int(self._dfunc), int(id_bits), C.byref(st)))
        self.last_stats = st


def fracture_model(grid: RegularGrid, params: FractureParameters):
    """CADScene::fractureModel (CADScene.cpp:624-691): seeds -> build -> erode | detectBoundaries(1).  Returns (seeds, stats)."""
    cap = (params._numSeeds + (max(params._biasSeeds, 0) if params._numImpacts > 0 else 0)) * 2 + params._numExtraSeeds
    seeds = np.zeros((cap, 4), dtype=np.uint32)
    n = C.c_uint32(0)
    st = VfFloodStats()
    check(grid._lib.vf_fracture_model(grid._h, C.byref(params._c), ptr(seeds), C.byref(n), C.byref(st)))
    return seeds[: n.value], st
