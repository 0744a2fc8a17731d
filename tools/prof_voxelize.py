"""Profiling helper: the SAT voxelizer (V2) on the synthetic vessel at a given maximum resolution; wall time per call by CUDA events."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import voxelfragmentml_b200 as vf
from voxelfragmentml_b200 import synth

ctx = vf.Context(0)
v, f = synth.vessel_mesh(0)
mn, mx = synth.mesh_aabb(v)
for res in [int(a) for a in sys.argv[1:]] or [256, 512]:
    dims = np.zeros(3, np.uint32)
    vf._capi.load().vf_dims_rule(mn.ctypes.data, mx.ctypes.data, res, dims.ctypes.data)
    dims = tuple(int(d) for d in dims)
    g = vf.RegularGrid(ctx, dims)
    g.setAABB(mn, mx, dims)
    ts = []
    for rep in range(6):
        ctx.timer_start(); g.fill(v, f); ts.append(ctx.timer_stop())
    occ = g.updateGrid()
    print("dims", dims, "triangles", len(f), "occupied", int((occ != 0).sum()), "voxelize ms", [round(t, 3) for t in ts], flush=True)
