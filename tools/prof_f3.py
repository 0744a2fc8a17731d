"""Profiling helper: one cfg4-style fragmentation (FLOOD CHEBYSHEV, 8 seeds + 16 extra seeds) of the 256-max vessel, for ncu launch lists."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import voxelfragmentml_b200 as vf
from voxelfragmentml_b200 import synth
ctx = vf.Context(0)
v, f = synth.vessel_mesh(0); mn, mx = synth.mesh_aabb(v)
dims = np.zeros(3, np.uint32); vf._capi.load().vf_dims_rule(mn.ctypes.data, mx.ctypes.data, 256, dims.ctypes.data); dims = tuple(int(d) for d in dims)
g = vf.RegularGrid(ctx, dims); g.setAABB(mn, mx, dims); g.fill(v, f)
occ = g.updateGrid()
fl = vf.FloodFracturer(); fl.setDistanceFunction(2)
for rep in range(2):
    g.updateSSBO(occ); ctx.initSeed(80); sd = vf.Seeder.make(g, 8, 16); ctx.synchronize()
    ctx.timer_start(); fl.build(g, sd); print(ctx.timer_stop(), "ms", flush=True)
