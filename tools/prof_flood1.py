import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import voxelfragmentml_b200 as vf
from voxelfragmentml_b200 import synth
ctx = vf.Context(0)
v, f = synth.vessel_mesh(0); mn, mx = synth.mesh_aabb(v)
dims = np.zeros(3, np.uint32); vf._capi.load().vf_dims_rule(mn.ctypes.data, mx.ctypes.data, 256, dims.ctypes.data); dims = tuple(int(d) for d in dims)
g = vf.RegularGrid(ctx, dims); g.setAABB(mn, mx, dims); g.fill(v, f)
ctx.initSeed(80); seeds = vf.Seeder.uniform(g, 16)
fl = vf.FloodFracturer(); fl.setDistanceFunction(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
ctx.timer_start(); fl.build(g, seeds); print(ctx.timer_stop(), "ms")
