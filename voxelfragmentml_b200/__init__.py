"""voxelfragmentml_b200 — host-side mirror of VoxelFragmentML's hot-path operators (RegularGrid, fracturer::Seeder,
fracturer::NaiveFracturer / FloodFracturer, FractureParameters) over the C ABI of libvoxfrag.so (hand-written CUDA,
sm_100a).  Importing the package loads the shared library and fails loudly when it is missing; nothing here computes
on the CPU."""
from . import _capi
from ._capi import SeederSearchError, VoxFragError

_capi.load()

from .api import (  # noqa: E402
    Context,
    DistanceFunction,
    ErosionType,
    ExportGrid,
    FloodFracturer,
    FractureAlgorithm,
    FractureParameters,
    NaiveFracturer,
    RandomUniformType,
    RegularGrid,
    Seeder,
    VOXEL_EMPTY,
    VOXEL_FREE,
    fracture_model,
)

from .dataset import FragmentationProcedure, generateDataset  # noqa: E402

__all__ = [
    "FragmentationProcedure", "generateDataset",
    "Context", "DistanceFunction", "ErosionType", "ExportGrid", "FloodFracturer", "FractureAlgorithm", "FractureParameters",
    "NaiveFracturer", "RandomUniformType", "RegularGrid", "Seeder", "SeederSearchError", "VoxFragError", "VOXEL_EMPTY",
    "VOXEL_FREE", "fracture_model",
]
