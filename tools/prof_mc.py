"""Times f2 (vf_marching_cubes: extraction + sort + fusion + smoothing) for every fragment of a fragmented vessel: python tools/prof_mc.py [maxvox]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import voxelfragmentml_b200 as vf
from voxelfragmentml_b200 import synth

maxvox = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ctx = vf.Context(0)
v, f = synth.vessel_mesh(0)
mn, mx = synth.mesh_aabb(v)
d = np.zeros(3, np.uint32)
ctx._lib.vf_dims_rule(vf._capi.ptr(np.float32(mn)), vf._capi.ptr(np.float32(mx)), maxvox, vf._capi.ptr(d))
dims = tuple(int(x) for x in d)
g = vf.RegularGrid(ctx, dims)
g.setAABB(mn, mx, dims)
for solid in (False, True):
    (g.fillSolid if solid else g.fill)(v, f)
    ctx.initSeed(80)
    p = vf.FractureParameters(_numSeeds=8, _numExtraSeeds=16)
    vf.fracture_model(g, p)
    counts, occ = g.countValues()
    labels = [int(i) for i in np.nonzero(counts)[0]]
    for rep in range(2):
        tot_v = tot_f = 0
        ctx.synchronize()
        ctx.timer_start()
        per = []
        for lab in labels:
            t0 = time.perf_counter()
            mv, mf = g.triangulateField(lab)
            per.append((lab, len(mf), round(1e3 * (time.perf_counter() - t0), 2)))
            tot_v += len(mv)
            tot_f += len(mf)
        ms = ctx.timer_stop()
    print("  per fragment (label, faces, ms):", per)
    print(f"{'solid' if solid else 'shell'} dims {dims}: {len(labels)} fragments, {occ} voxels -> {tot_v} vertices, {tot_f} faces, {ms:.2f} ms for all fragments "
          f"({ms / len(labels):.2f} ms per fragment, download included)")
