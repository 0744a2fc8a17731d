"""Profiling helper: flood timings on (a) the SAT-voxelized synthetic vessel at 256 (cfg2), (b) dense n^3."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import voxelfragmentml_b200 as vf
from voxelfragmentml_b200 import synth

ctx = vf.Context(0)
res = int(sys.argv[1]) if len(sys.argv) > 1 else 256
if len(sys.argv) > 3:
    ctx.setFloodFront(int(sys.argv[3]))  # 0 = tiles only; default: thin-front solver first (8192 pending cells)
    print("flood front limit", int(sys.argv[3]), flush=True)
v, f = synth.vessel_mesh(0)
mn, mx = synth.mesh_aabb(v)
dims = np.zeros(3, np.uint32)
vf._capi.load().vf_dims_rule(mn.ctypes.data, mx.ctypes.data, res, dims.ctypes.data)
dims = tuple(int(d) for d in dims)
g = vf.RegularGrid(ctx, dims)
g.setAABB(mn, mx, dims)
for rep in range(3):
    ctx.timer_start(); g.fill(v, f); t_vox = ctx.timer_stop()
occ = g.updateGrid()
print("dims", dims, "occupied", int((occ != 0).sum()), f"voxelize {t_vox:.3f} ms", flush=True)
ctx.initSeed(80)
seeds = vf.Seeder.uniform(g, 16)
for name, df, extra in [("manhattan16", 1, 0), ("chebyshev16", 2, 0), ("cheb 8+16 extra", 2, 16)]:
    fl = vf.FloodFracturer(); fl.setDistanceFunction(df)
    sd = seeds if not extra else None
    if extra:
        ctx.initSeed(80); sd = vf.Seeder.make(g, 8, extra)
    for rep in range(3):
        g.updateSSBO(occ); ctx.synchronize()
        ctx.timer_start(); fl.build(g, sd); ms = ctx.timer_stop()
    st = fl.last_stats
    print(f"{name}: {ms:.3f} ms front levels {st.front_levels} rounds {st.tile_rounds} visits {st.tile_visits} maxdist {st.max_dist} disjoint {st.disjoint_rounds} freed {st.freed_voxels}", flush=True)
n = int(sys.argv[2]) if len(sys.argv) > 2 else 256
g2 = vf.RegularGrid(ctx, (n, n, n))
rs = np.random.RandomState(1); pts = rs.randint(0, n, size=(16, 3)); pts = pts[np.lexsort((pts[:, 2], pts[:, 1], pts[:, 0]))]
sd = np.concatenate([pts, np.arange(2, 18)[:, None]], 1).astype(np.uint32)
for df in (1, 2):
    fl = vf.FloodFracturer(); fl.setDistanceFunction(df)
    for rep in range(2):
        g2.fillValue(1); ctx.synchronize(); ctx.timer_start(); fl.build(g2, sd); ms = ctx.timer_stop()
    st = fl.last_stats
    print(f"dense {n}^3 df={df}: {ms:.3f} ms front levels {st.front_levels} rounds {st.tile_rounds} visits {st.tile_visits} maxdist {st.max_dist}", flush=True)
