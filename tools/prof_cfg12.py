"""Times BASELINE configs 1 and 2 end to end on the device (mesh upload -> voxelize -> seeds -> fragment -> detectBoundaries -> histogram+undoMask):
cfg1 = 128-max vessel, NAIVE EUCLIDEAN, 8 seeds; cfg2 = 256-max vessel, FLOOD MANHATTAN, 16 seeds.  python tools/prof_cfg12.py"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import voxelfragmentml_b200 as vf
from voxelfragmentml_b200 import synth

ctx = vf.Context(0)
v, f = synth.vessel_mesh(0)
mn, mx = synth.mesh_aabb(v)
for name, maxvox, kw in (("cfg1", 128, dict(_fractureAlgorithm=vf.FractureAlgorithm.NAIVE, _distanceFunction=vf.DistanceFunction.EUCLIDEAN, _numSeeds=8, _numExtraSeeds=0)),
                         ("cfg2", 256, dict(_fractureAlgorithm=vf.FractureAlgorithm.FLOOD, _distanceFunction=vf.DistanceFunction.MANHATTAN, _numSeeds=16, _numExtraSeeds=0))):
    d = np.zeros(3, np.uint32)
    ctx._lib.vf_dims_rule(vf._capi.ptr(np.float32(mn)), vf._capi.ptr(np.float32(mx)), maxvox, vf._capi.ptr(d))
    dims = tuple(int(x) for x in d)
    g = vf.RegularGrid(ctx, dims)
    ctx.reserve(dims)
    p = vf.FractureParameters(**kw)
    parts = {}

    def step(record):
        t = {}

        def timed(key, fn):
            ctx.timer_start()
            r = fn()
            t[key] = ctx.timer_stop()
            return r

        ctx.initSeed(80)
        timed("setAABB+voxelize", lambda: (g.setAABB(mn, mx, dims), g.fill(v, f)))
        timed("seeds+fragment+detectBoundaries", lambda: vf.fracture_model(g, p))
        counts, occ = timed("histogram+undoMask", lambda: g.countValuesUndoMask())
        if record:
            for k, val in t.items():
                parts.setdefault(k, []).append(val)
        return occ

    for i in range(13):
        t0 = time.perf_counter()
        occ = step(i >= 3)
        wall = time.perf_counter() - t0
    med = {k: float(np.median(val)) for k, val in parts.items()}
    tot = sum(med.values())
    n = dims[0] * dims[1] * dims[2]
    print(f"{name}: dims {dims} ({n} cells, {occ} occupied), {len(f)} triangles: " + ", ".join(f"{k} {val:.3f} ms" for k, val in med.items()) +
          f" | total {tot:.3f} ms = {n / tot / 1e6:.2f} Gvoxels/s of grid, {1e3 / tot:.0f} models/s serial (last step wall {wall * 1e3:.3f} ms)")
    g.close()
