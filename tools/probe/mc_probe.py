import sys, time, ctypes as C
import numpy as np
sys.path.insert(0, ".")
import voxelfragmentml_b200 as vf
from voxelfragmentml_b200 import synth, _capi
ctx = vf.Context(0)
v, f = synth.vessel_mesh(0)
mn, mx = synth.mesh_aabb(v)
dims = (176, 256, 176)
g = vf.RegularGrid(ctx, dims); g.setAABB(mn, mx, dims); g.fillSolid(v, f)
ctx.initSeed(80)
vf.fracture_model(g, vf.FractureParameters(_numSeeds=8, _numExtraSeeds=16))
counts, occ = g.countValues()
labels = [int(i) for i in np.nonzero(counts)[0]]
lib = g._lib
for rep in range(2):
    for lab in labels[:4]:
        t0 = time.perf_counter()
        h = C.c_void_p(); mp = _capi.VfMcParams(0.048, 0.2, 0.048, 0.9, 1)
        lib.vf_marching_cubes(g._h, lab, C.byref(mp), C.byref(h)); t1 = time.perf_counter()
        ctx.synchronize(); t2 = time.perf_counter()
        nv, nf = C.c_uint32(0), C.c_uint32(0); lib.vf_mesh_counts(h, C.byref(nv), C.byref(nf))
        vv = np.zeros((nv.value, 4), np.float32); ff = np.zeros((nf.value, 4), np.uint32); t3 = time.perf_counter()
        lib.vf_mesh_download(h, vv.ctypes.data, ff.ctypes.data); t4 = time.perf_counter()
        lib.vf_mesh_destroy(h); t5 = time.perf_counter()
        print(f"rep {rep} label {lab}: nv {nv.value} nf {nf.value} | call {1e3*(t1-t0):.2f} sync {1e3*(t2-t1):.2f} alloc {1e3*(t3-t2):.2f} download {1e3*(t4-t3):.2f} destroy {1e3*(t5-t4):.2f} ms")
