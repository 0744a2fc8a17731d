"""Regenerates tests/golden/qstack_golden.json: sha256 + length of the `.qstack` bytes the REFERENCE'S OWN QuadStack / GStack
headers (compiled in place into oracle/_ref/libvf_ref_qstack.so, driven with exportQuadStack's call sequence, RegularGrid.cpp:716-725)
write for the deterministic grids of tests/vox_cases.py::all_qstack_cases.  QuadStack::loadCube keeps a function-static matrix that
only grows, so every export loads a fresh copy of the library (= "first export of a process").
Run in the build container only (needs /root/reference for `make -C oracle/ref_shim`)."""
import ctypes as C
import hashlib
import json
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from vox_cases import all_qstack_cases  # noqa: E402

REF_SO = os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle", "_ref", "libvf_ref_qstack.so")


def ref_qstack_bytes(grid, tmp):
    """bytes written by the reference's QuadStack for `grid` (fresh library image per call)"""
    n = len(os.listdir(tmp))
    so = os.path.join(tmp, f"q{n}.so")
    shutil.copy(REF_SO, so)
    lib = C.CDLL(so)
    lib.ref_export_qstack.restype = C.c_int
    lib.ref_export_qstack.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p]
    g = np.ascontiguousarray(grid, np.uint16)
    d = np.asarray(g.shape, np.uint32)
    path = os.path.join(tmp, f"g{n}.qstack")
    assert lib.ref_export_qstack(g.ctypes.data, d.ctypes.data, path.encode()) == 0
    return open(path, "rb").read()


if __name__ == "__main__":
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for name, grid in all_qstack_cases():
            data = ref_qstack_bytes(grid, tmp)
            out[name] = {"bytes": len(data), "sha256": hashlib.sha256(data).hexdigest()}
    json.dump(out, open(os.path.join(HERE, "qstack_golden.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))
