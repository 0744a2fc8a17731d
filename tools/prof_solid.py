"""Times V1 (vf_voxelize_solid) and V2 (vf_voxelize) on the synthetic vessel mesh: python tools/prof_solid.py [maxvox ...]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import voxelfragmentml_b200 as vf
from voxelfragmentml_b200 import synth

ctx = vf.Context(0)
v, f = synth.vessel_mesh(0)
mn, mx = synth.mesh_aabb(v)
for maxvox in [int(a) for a in sys.argv[1:]] or [128, 256, 512]:
    d = np.zeros(3, np.uint32)
    ctx._lib.vf_dims_rule(vf._capi.ptr(np.float32(mn)), vf._capi.ptr(np.float32(mx)), maxvox, vf._capi.ptr(d))
    dims = tuple(int(x) for x in d)
    g = vf.RegularGrid(ctx, dims)
    g.setAABB(mn, mx, dims)
    for name, fn in (("solid", lambda: g.fillSolid(v, f)), ("sat", lambda: g.fill(v, f))):
        for _ in range(3):
            fn()
        ctx.synchronize()
        ts = []
        for _ in range(10):
            ctx.timer_start()
            fn()
            ts.append(ctx.timer_stop())
        t0 = time.perf_counter()
        for _ in range(10):
            fn()
        ctx.synchronize()
        wall = (time.perf_counter() - t0) / 10 * 1e3
        occ = int((g.updateGrid() != 0).sum())
        print(f"{name} dims {dims} tris {len(f)}: device {min(ts):.3f} ms (median {sorted(ts)[5]:.3f}), wall/call {wall:.3f} ms, occupied {occ}")
    g.close()
