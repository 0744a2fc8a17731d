"""Design aid: per-phase cycles of a flood tile visit (library built with VF_NVCC_EXTRA=-DVF_FLOOD_TIMING python build_lib.py --force)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import voxelfragmentml_b200 as vf
from voxelfragmentml_b200 import synth
lib = vf._capi.load()
ctx = vf.Context(0)
v, f = synth.vessel_mesh(0); mn, mx = synth.mesh_aabb(v)
dims = np.zeros(3, np.uint32); lib.vf_dims_rule(mn.ctypes.data, mx.ctypes.data, 256, dims.ctypes.data); dims = tuple(int(d) for d in dims)
g = vf.RegularGrid(ctx, dims); g.setAABB(mn, mx, dims); g.fill(v, f); occ = g.updateGrid()
ctx.initSeed(80); seeds = vf.Seeder.uniform(g, 16)
out = (C.c_ulonglong * 8)()
for df in (1, 2):
    fl = vf.FloodFracturer(); fl.setDistanceFunction(df)
    for rep in range(2):
        g.updateSSBO(occ); ctx.synchronize(); lib.vf_debug_flood_cycles(out, 1)
        ctx.timer_start(); fl.build(g, seeds); ms = ctx.timer_stop()
    lib.vf_debug_flood_cycles(out, 1)
    nv = max(1, out[4])
    print(f"df {df}: {ms:.3f} ms, visits {out[4]}, steps/visit {out[5]/nv:.1f}, cycles/visit: load {out[0]/nv:.0f} masks {out[1]/nv:.0f} relax {out[2]/nv:.0f} ({out[2]/max(1,out[5]):.0f}/step) store+wake {out[3]/nv:.0f}; entry step {out[6]/nv:.0f}, later steps {(out[2]-out[6])/max(1,out[5]-out[4]):.0f} each, of which warp 0 waits at the barrier {out[7]/max(1,out[5]):.0f} per step")
