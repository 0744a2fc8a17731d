"""Host-side mirror of the reference's dataset generator: struct FragmentationProcedure (SRC/Graphics/Core/FragmentationProcedure.h)
and CADScene::generateDataset (SRC/Graphics/Application/CADScene.cpp:209-507), voxel path only, over libvoxfrag's native driver
(csrc/dataset.cpp).  Nothing here computes: the loop, the file writers and the .obj reader are native."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from ._capi import VfDatasetStats, VfProcedure, check, ptr
from .api import Context, FractureParameters, RegularGrid


class FragmentationProcedure:
    """Same field names (leading underscore) and defaults as the reference's struct; `_fractureParameters` is a FractureParameters."""

    def __init__(self, **kw):
        self._c = VfProcedure()
        _capi.load().vf_procedure_default(C.byref(self._c))
        self._fractureParameters = FractureParameters()
        C.memmove(C.byref(self._fractureParameters._c), C.byref(self._c.fractureParameters), C.sizeof(self._c.fractureParameters))
        self._fragmentInterval = (2, 10)
        self._iterationInterval = (25, 15)
        self._maxFragmentsModel = 1000
        self._exportGrid = True
        self._startVessel = ""
        self._searchExtension = ".obj"
        self._exportMesh = False          # fragment meshes as .binm + mesh metadata rows (the reference's _exportMesh with _targetTriangles empty)
        self._solidVoxelization = False   # extension: Tetravoxelizer occupancy instead of the SAT surface occupancy
        self._writerThreads = 2           # extension: asynchronous file writers
        for k, v in kw.items():
            if not hasattr(self, k):
                raise AttributeError(k)
            setattr(self, k, v)

    def _struct(self) -> VfProcedure:
        c = self._c
        C.memmove(C.byref(c.fractureParameters), C.byref(self._fractureParameters._c), C.sizeof(c.fractureParameters))
        c.fragmentInterval[0], c.fragmentInterval[1] = (int(v) for v in self._fragmentInterval)
        c.iterationInterval[0], c.iterationInterval[1] = (int(v) for v in self._iterationInterval)
        c.maxFragmentsModel = int(self._maxFragmentsModel)
        c.exportGrid = int(bool(self._exportGrid))
        c.solidVoxelization = int(bool(self._solidVoxelization))
        c.exportMesh = int(bool(self._exportMesh))
        c.writerThreads = int(self._writerThreads)
        return c

    def numIterations(self, numFragments: int) -> int:
        return int(_capi.load().vf_dataset_iterations(C.byref(self._struct()), int(numFragments)))


def dataset_dims(aabb_min, aabb_max, voxelPerMetricUnit: int, clampVoxelMetricUnit: int):
    """voxelization size of a model in dataset mode (CADScene.cpp:262-273)"""
    d = np.zeros(3, np.uint32)
    mn, mx = np.ascontiguousarray(aabb_min, np.float32), np.ascontiguousarray(aabb_max, np.float32)
    _capi.load().vf_dataset_dims_rule(ptr(mn), ptr(mx), int(voxelPerMetricUnit), int(clampVoxelMetricUnit), ptr(d))
    return tuple(int(v) for v in d)


def _stats_dict(st: VfDatasetStats) -> dict:
    return {name: getattr(st, name) for name, _ in st._fields_}


def dataset_grid(ctx: Context, procedure: FragmentationProcedure) -> RegularGrid:
    """the one grid of a dataset run, allocated at the clamp size (CADScene::allocateMemoryDataset, CADScene.cpp:529-543)"""
    clamp = int(procedure._fractureParameters._clampVoxelMetricUnit)
    dims = (clamp + 3, clamp, clamp + 3)  # x and z may be rounded up to a multiple of 4
    ctx.reserve(dims)
    return RegularGrid(ctx, dims)


def generate_model(grid: RegularGrid, procedure: FragmentationProcedure, modelName: str, vertices, faces, destinationFolder: str, stats=None) -> dict:
    """the body of generateDataset's model loop for one loaded model; continues the context's RNG stream"""
    v = np.ascontiguousarray(vertices, np.float32)
    f = np.ascontiguousarray(faces, np.uint32)
    st = stats if stats is not None else VfDatasetStats()
    check(grid._lib.vf_dataset_model(grid._h, C.byref(procedure._struct()), modelName.encode(), ptr(v), len(v), ptr(f), len(f),
                                     destinationFolder.encode(), C.byref(st)))
    return _stats_dict(st)


def generateDataset(ctx: Context, procedure: FragmentationProcedure, folder: str, extension: str, destinationFolder: str) -> dict:
    """CADScene::generateDataset(procedure, folder, extension, destinationFolder)"""
    st = VfDatasetStats()
    check(ctx._lib.vf_dataset_generate(ctx._h, C.byref(procedure._struct()), folder.encode(), extension.encode(),
                                       procedure._startVessel.encode(), destinationFolder.encode(), C.byref(st)))
    return _stats_dict(st)


def load_obj(path: str):
    """minimal Wavefront reader + CADModel::load normalisation (native); returns (vertices float32[nv][3], faces uint32[nf][3])"""
    lib = _capi.load()
    pv, pf = C.POINTER(C.c_float)(), C.POINTER(C.c_uint32)()
    nv, nf = C.c_uint32(0), C.c_uint32(0)
    check(lib.vf_load_obj(path.encode(), C.byref(pv), C.byref(nv), C.byref(pf), C.byref(nf)))
    try:
        v = np.ctypeslib.as_array(pv, shape=(nv.value, 3)).copy()
        f = np.ctypeslib.as_array(pf, shape=(nf.value, 3)).copy()
    finally:
        lib.vf_free_host(pv), lib.vf_free_host(pf)
    return v, f
