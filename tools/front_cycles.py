"""Design aid: where a BFS level of the thin-front flood solver spends its cycles (library built with VF_NVCC_EXTRA=-DVF_FLOOD_TIMING python
build_lib.py --force; VF_FRONT_CLUSTER=16|8|4|2|1 caps the cluster size in that build)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import voxelfragmentml_b200 as vf
from voxelfragmentml_b200 import synth
lib = vf._capi.load()
ctx = vf.Context(0)
ctx.setFloodFront(65536)
v, f = synth.vessel_mesh(0); mn, mx = synth.mesh_aabb(v)
dims = np.zeros(3, np.uint32); lib.vf_dims_rule(mn.ctypes.data, mx.ctypes.data, 256, dims.ctypes.data); dims = tuple(int(d) for d in dims)
g = vf.RegularGrid(ctx, dims); g.setAABB(mn, mx, dims); g.fill(v, f); occ = g.updateGrid()
ctx.initSeed(80); seeds = vf.Seeder.uniform(g, 16)
out = (C.c_ulonglong * 12)()
for df in (1, 2):
    fl = vf.FloodFracturer(); fl.setDistanceFunction(df)
    for rep in range(3):
        g.updateSSBO(occ); ctx.synchronize(); lib.vf_debug_front_cycles(out, 1)
        ctx.timer_start(); fl.build(g, seeds); ms = ctx.timer_stop()
    lib.vf_debug_front_cycles(out, 1)
    nl = max(1, out[7])
    print(f"cluster {out[8]} df {df}: {ms:.3f} ms, steps {out[7]} in {out[11]} barrier intervals (max_dist {fl.last_stats.max_dist}), pending/interval {out[9]/max(1,out[11]):.0f}, passes/step {out[10]/nl:.2f}, "
          f"cycles/step: set-up {out[0]/nl:.0f} key loads {out[1]/nl:.0f} claims {out[2]/nl:.0f} pushes {out[3]/nl:.0f} step end {out[4]/nl:.0f} barrier + merge {out[5]/nl:.0f} total {sum(out[:7])/nl:.0f}", flush=True)
