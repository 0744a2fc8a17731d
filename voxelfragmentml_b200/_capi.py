"""ctypes binding of libvoxfrag.so (include/voxfrag.h).  No compute happens in Python and there is no fallback:
if the shared library is missing, import fails loudly."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libvoxfrag.so")

HISTOGRAM_BINS = 32768


class VfParams(C.Structure):
    """struct vf_params == the reference's FractureParameters fields used by the path (FractureParameters.h:43-85)."""

    _fields_ = [
        ("biasFocus", C.c_int32), ("biasSeeds", C.c_int32), ("clampVoxelMetricUnit", C.c_int32), ("erode", C.c_int32),
        ("erosionConvolution", C.c_int32), ("erosionIterations", C.c_int32), ("erosionProbability", C.c_float),
        ("erosionSize", C.c_int32), ("erosionThreshold", C.c_float), ("fractureAlgorithm", C.c_int32),
        ("distanceFunction", C.c_int32), ("launchGPU", C.c_int32), ("mergeSeedsDistanceFunction", C.c_int32),
        ("neighbourhoodType", C.c_int32), ("numExtraSeeds", C.c_int32), ("numImpacts", C.c_int32), ("numSeeds", C.c_int32),
        ("removeIsolatedRegions", C.c_int32), ("seed", C.c_int32), ("seedingRandom", C.c_int32),
        ("voxelPerMetricUnit", C.c_int32), ("voxelizationSize", C.c_int32 * 3), ("exportGridExtension", C.c_int32),
        ("floodIdBits", C.c_int32), ("erodeBoundaryMode", C.c_int32),
    ]


class VfMcParams(C.Structure):
    _fields_ = [("boundaryMCIterations", C.c_float), ("boundaryMCWeight", C.c_float), ("nonBoundaryMCIterations", C.c_float),
                ("nonBoundaryMCWeight", C.c_float), ("marchingCubesSubdivisions", C.c_int32)]


class VfProcedure(C.Structure):
    """struct vf_procedure == FragmentationProcedure's voxel-path fields (FragmentationProcedure.h:6-60)."""

    _fields_ = [("fractureParameters", VfParams), ("fragmentInterval", C.c_int32 * 2), ("iterationInterval", C.c_int32 * 2),
                ("maxFragmentsModel", C.c_uint64), ("exportGrid", C.c_int32), ("solidVoxelization", C.c_int32), ("exportMesh", C.c_int32),
                ("writerThreads", C.c_int32)]


class VfDatasetStats(C.Structure):
    _fields_ = [("models", C.c_uint64), ("fragmentations", C.c_uint64), ("fragments", C.c_uint64), ("files", C.c_uint64),
                ("bytes_written", C.c_uint64), ("bytes_downloaded", C.c_uint64), ("voxels", C.c_uint64),
                ("seconds_voxelize", C.c_double), ("seconds_fracture", C.c_double), ("seconds_export", C.c_double)]


class VfFloodStats(C.Structure):
    _fields_ = [("tile_rounds", C.c_uint32), ("tile_visits", C.c_uint32), ("disjoint_rounds", C.c_uint32),
                ("freed_voxels", C.c_uint32), ("max_dist", C.c_uint32), ("front_levels", C.c_uint32)]


_vp = C.c_void_p
_u32 = C.c_uint32
_f32p = C.POINTER(C.c_float)

# name -> (restype, argtypes).  Kept in one table so tests can check that every symbol include/voxfrag.h declares is exported.
SIGNATURES = {
    "vf_last_error": (C.c_char_p, []),
    "vf_version": (C.c_char_p, []),
    "vf_device_count": (C.c_int, []),
    "vf_params_default": (None, [C.POINTER(VfParams)]),
    "vf_ctx_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "vf_ctx_create_on_stream": (C.c_int, [C.c_int, _vp, C.POINTER(_vp)]),
    "vf_ctx_destroy": (None, [_vp]),
    "vf_ctx_reserve": (C.c_int, [_vp, _u32, _u32, _u32]),
    "vf_ctx_synchronize": (C.c_int, [_vp]),
    "vf_ctx_set_blocking_sync": (C.c_int, [_vp, C.c_int]),
    "vf_ctx_set_flood_levels": (C.c_int, [_vp, _u32]),
    "vf_ctx_set_flood_front": (C.c_int, [_vp, _u32]),
    "vf_ctx_set_c1_mode": (C.c_int, [_vp, C.c_int]),
    "vf_ctx_set_flood_mode": (C.c_int, [_vp, C.c_int]),
    "vf_ctx_stream": (_vp, [_vp]),
    "vf_ctx_kernel_launches": (C.c_uint64, [_vp]),
    "vf_ctx_host_waits": (C.c_uint64, [_vp]),
    "vf_ctx_timer_start": (C.c_int, [_vp]),
    "vf_ctx_timer_stop": (C.c_int, [_vp, _f32p]),
    "vf_rng_seed": (C.c_int, [_vp, _u32]),
    "vf_crand_seed": (C.c_int, [_vp, _u32]),
    "vf_crand_next": (C.c_int, [_vp]),
    "vf_seed_near": (C.c_int, [_vp, _vp, _u32, _u32, _u32, _u32, _vp, _u32, C.POINTER(_u32)]),
    "vf_rng_uniform": (C.c_float, [_vp]),
    "vf_rng_raw": (C.c_uint32, [_vp]),
    "vf_fill_noise": (C.c_int, [_vp, _vp, _u32]),
    "vf_grid_create": (C.c_int, [_vp, _u32, _u32, _u32, C.POINTER(_vp)]),
    "vf_grid_wrap": (C.c_int, [_vp, _vp, _u32, _u32, _u32, C.POINTER(_vp)]),
    "vf_grid_destroy": (None, [_vp]),
    "vf_grid_set_aabb": (C.c_int, [_vp, _vp, _vp, _u32, _u32, _u32]),
    "vf_grid_dims": (C.c_int, [_vp, _vp]),
    "vf_grid_device_ptr": (_vp, [_vp]),
    "vf_grid_upload": (C.c_int, [_vp, _vp]),
    "vf_grid_download": (C.c_int, [_vp, _vp]),
    "vf_grid_upload_async": (C.c_int, [_vp, _vp]),
    "vf_grid_upload_bits": (C.c_int, [_vp, _vp]),
    "vf_grid_download_async": (C.c_int, [_vp, _vp]),
    "vf_grid_fill": (C.c_int, [_vp, C.c_uint16]),
    "vf_dims_rule": (None, [_vp, _vp, _u32, _vp]),
    "vf_voxelize": (C.c_int, [_vp, _vp, _u32, _vp, _u32]),
    "vf_voxelize_solid": (C.c_int, [_vp, _vp, _u32, _vp, _u32, _vp]),
    "vf_seed_uniform": (C.c_int, [_vp, _u32, C.c_int, C.c_int, _vp, C.POINTER(_u32)]),
    "vf_merge_seeds": (C.c_int, [_vp, _u32, _vp, _u32, C.c_int]),
    "vf_make_seeds": (C.c_int, [_vp, _u32, _u32, C.c_int, C.c_int, _vp, _u32, C.POINTER(_u32)]),
    "vf_fracture_naive": (C.c_int, [_vp, _vp, _u32, C.c_int]),
    "vf_fracture_naive_slab": (C.c_int, [_vp, _vp, _u32, C.c_int, _u32, _u32]),
    "vf_fracture_flood": (C.c_int, [_vp, _vp, _u32, C.c_int, C.c_int, C.POINTER(VfFloodStats)]),
    "vf_flood_slab_init": (C.c_int, [_vp, _vp, _vp, _u32, C.c_int, C.c_int, C.c_int, C.POINTER(_vp)]),
    "vf_flood_slab_relax": (C.c_int, [_vp, C.POINTER(C.c_uint64)]),
    "vf_flood_slab_boundary_ptr": (_vp, [_vp, C.c_int]),
    "vf_flood_slab_ingest": (C.c_int, [_vp, C.c_int, _vp, C.POINTER(C.c_uint64)]),
    "vf_flood_slab_finalize": (C.c_int, [_vp, _vp, _u32, C.POINTER(_u32)]),
    "vf_flood_slab_destroy": (None, [_vp]),
    "vf_nccl_unique_id": (C.c_int, [_vp]),
    "vf_nccl_comm_create": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.POINTER(_vp)]),
    "vf_nccl_comm_destroy": (None, [_vp]),
    "vf_flood_slab_run": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.POINTER(_u32), C.POINTER(C.c_uint64)]),
    "vf_remove_isolated_regions": (C.c_int, [_vp, _vp, _u32]),
    "vf_detect_boundaries": (C.c_int, [_vp, C.c_int]),
    "vf_erode": (C.c_int, [_vp, C.c_int, _u32, _u32, C.c_float, C.c_float, _vp, _u32, C.c_int]),
    "vf_erode_pass": (C.c_int, [_vp, C.c_int, _u32, C.c_float, C.c_float, _vp, _u32, C.c_int, C.c_uint64]),
    "vf_remove_isolated_regions_grid": (C.c_int, [_vp]),
    "vf_undo_mask": (C.c_int, [_vp]),
    "vf_reset_filling": (C.c_int, [_vp]),
    "vf_homogenize": (C.c_int, [_vp]),
    "vf_histogram": (C.c_int, [_vp, _vp, C.POINTER(C.c_uint64)]),
    "vf_histogram_undo_mask": (C.c_int, [_vp, _vp, C.POINTER(C.c_uint64)]),
    "vf_export": (C.c_int, [_vp, C.c_char_p, C.c_int, C.c_int]),
    "vf_mc_params_default": (None, [C.POINTER(VfMcParams)]),
    "vf_marching_cubes": (C.c_int, [_vp, _u32, C.POINTER(VfMcParams), C.POINTER(_vp)]),
    "vf_mesh_counts": (C.c_int, [_vp, C.POINTER(_u32), C.POINTER(_u32)]),
    "vf_mesh_download": (C.c_int, [_vp, _vp, _vp]),
    "vf_mesh_destroy": (None, [_vp]),
    "vf_procedure_default": (None, [C.POINTER(VfProcedure)]),
    "vf_dataset_dims_rule": (None, [_vp, _vp, C.c_int32, C.c_int32, _vp]),
    "vf_dataset_iterations": (C.c_int32, [C.POINTER(VfProcedure), C.c_int32]),
    "vf_dataset_model": (C.c_int, [_vp, C.POINTER(VfProcedure), C.c_char_p, _vp, _u32, _vp, _u32, C.c_char_p, C.POINTER(VfDatasetStats)]),
    "vf_dataset_generate": (C.c_int, [_vp, C.POINTER(VfProcedure), C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(VfDatasetStats)]),
    "vf_load_obj": (C.c_int, [C.c_char_p, C.POINTER(C.POINTER(C.c_float)), C.POINTER(_u32), C.POINTER(C.POINTER(_u32)), C.POINTER(_u32)]),
    "vf_free_host": (None, [_vp]),
    "vf_grid_encode_rle": (C.c_int, [_vp, _vp, C.c_uint64, _vp]),
    "vf_encode_rle": (C.c_uint64, [_vp, _vp, _vp, C.c_uint64]),
    "vf_encode_bing_squared": (C.c_uint64, [_vp, _vp, _vp, C.c_uint64]),
    "vf_encode_vox": (C.c_uint64, [_vp, _vp, C.c_int, _vp, C.c_uint64]),
    "vf_encode_qstack": (C.c_uint64, [_vp, _vp, _vp, C.c_uint64]),
    "vf_synth_solid_vessel": (C.c_int, [_vp, C.c_int, _u32, C.c_float, C.c_float, C.c_float]),
    "vf_fracture_model": (C.c_int, [_vp, C.POINTER(VfParams), _vp, C.POINTER(_u32), C.POINTER(VfFloodStats)]),
}

STATUS_NAMES = {0: "VF_OK", 1: "VF_ERR_INVALID_ARGUMENT", 2: "VF_ERR_SEEDER_EXHAUSTED", 3: "VF_ERR_INVALID_DISTANCE",
                4: "VF_ERR_CAPACITY", 5: "VF_ERR_CUDA", 6: "VF_ERR_IO", 7: "VF_ERR_UNSUPPORTED"}


class VoxFragError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"{STATUS_NAMES.get(status, status)}: {message}")
        self.status = status


class SeederSearchError(VoxFragError):
    """fracturer::Seeder::SeederSearchError (SRC/Fracturer/Seeder.h:17-22)."""


_lib = None


def load():
    """Load libvoxfrag.so; raises (never falls back) when the CUDA extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python build_lib.py` "
            "(nvcc, sm_100a).  voxelfragmentml_b200 has no CPU or PyTorch fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the ABI and this table diverge
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int):
    if status != 0:
        msg = load().vf_last_error().decode("utf-8", "replace")
        if status == 2:
            raise SeederSearchError(status, msg)
        raise VoxFragError(status, msg)


def ptr(a) -> int:
    """host numpy array / torch tensor / int address -> raw address"""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"]
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    raise TypeError(type(a))
