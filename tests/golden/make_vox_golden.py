"""Regenerates tests/golden/vox_golden.json: sha256 + length of the `.vox` bytes the REFERENCE'S OWN MagicaVoxel writer
(oracle/_ref/libvf_ref.so = Libraries/MagicaVoxel_File_Writer/VoxWriter.cpp compiled in place, driven with exportVox's AddVoxel
sequence) produces for the deterministic grids of tests/test_oracle_vs_ref.py::vox_cases plus the labelled AL_12B fixture.
Run in the build container only (needs /root/reference for `make -C oracle/ref_shim`)."""
import ctypes as C
import hashlib
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from vox_cases import all_vox_cases as all_cases  # noqa: E402

REF_SO = os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle", "_ref", "libvf_ref.so")


if __name__ == "__main__":
    ref = C.CDLL(REF_SO)
    u16 = np.ctypeslib.ndpointer(np.uint16, flags="C_CONTIGUOUS")
    u32 = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
    ref.ref_export_vox.argtypes = [u16, u32, C.c_int, C.c_char_p]
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for name, grid in all_cases():
            for squared in (0, 1):
                path = os.path.join(tmp, "g.vox")
                ref.ref_export_vox(np.ascontiguousarray(grid), np.asarray(grid.shape, np.uint32), squared, path.encode())
                data = open(path, "rb").read()
                out[f"{name}/{'squared' if squared else 'tight'}"] = {"bytes": len(data), "sha256": hashlib.sha256(data).hexdigest()}
    json.dump(out, open(os.path.join(HERE, "vox_golden.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))
