"""Design aid: per-mesh time of the cfg4 loop (one context, one job) over distinct synthetic shapes: finds shapes that cost far more than the others."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import voxelfragmentml_b200 as vf
from voxelfragmentml_b200 import synth
lib = vf._capi.load()
ctx = vf.Context(0); ctx.setFloodLevels(8); ctx.setFloodMode(0)
grid = vf.RegularGrid(ctx, (256, 256, 256)); ctx.reserve((256, 256, 256))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 48
rows = []
for rep in range(2):
    for m in range(n):
        v, f = synth.vessel_mesh(m)
        mn, mx = synth.mesh_aabb(v)
        dims = np.zeros(3, np.uint32); lib.vf_dims_rule(mn.ctypes.data, mx.ctypes.data, 256, dims.ctypes.data)
        t0 = time.perf_counter()
        grid.setAABB(mn, mx, tuple(int(d) for d in dims)); grid.fill(v, f); ctx.synchronize()
        t1 = time.perf_counter()
        ctx.initSeed(80 + m)
        w0, l0 = ctx.host_waits, ctx.kernel_launches
        occs = []
        for k in range(10):
            nf = 2 + (k % 9)
            p = vf.FractureParameters(_numSeeds=nf, _numExtraSeeds=2 * nf)
            grid.homogenize(); vf.fracture_model(grid, p)
            counts, occ = grid.countValuesUndoMask(); occs.append(occ)
        t2 = time.perf_counter()
        if rep == 1:
            rows.append((m, tuple(int(d) for d in dims), len(f), occs[0], (t1 - t0) * 1e3, (t2 - t1) * 1e3, ctx.host_waits - w0, ctx.kernel_launches - l0))
for r in rows:
    print("mesh %3d dims %s tris %d occ %8d  voxelize %6.2f ms  10 fragmentations %7.2f ms  waits %4d launches %5d" % r)
tot = sum(r[5] for r in rows)
print("mean per mesh %.2f ms, max %.2f ms, min %.2f ms" % (tot / len(rows), max(r[5] for r in rows), min(r[5] for r in rows)))
