"""Design aid: flood latency against the width of a round's distance window (vf_ctx_set_flood_levels), cooperative round loop."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import voxelfragmentml_b200 as vf
from voxelfragmentml_b200 import synth
ctx = vf.Context(0)
v, f = synth.vessel_mesh(0); mn, mx = synth.mesh_aabb(v)
dims = np.zeros(3, np.uint32); vf._capi.load().vf_dims_rule(mn.ctypes.data, mx.ctypes.data, 256, dims.ctypes.data); dims = tuple(int(d) for d in dims)
g = vf.RegularGrid(ctx, dims); g.setAABB(mn, mx, dims); g.fill(v, f); occ = g.updateGrid()
ctx.initSeed(80); seeds = vf.Seeder.uniform(g, 16)
ctx.initSeed(80); xs = vf.Seeder.make(g, 8, 16)
n = 256
g2 = vf.RegularGrid(ctx, (n, n, n))
rs = np.random.RandomState(1); pts = rs.randint(0, n, size=(16, 3)); pts = pts[np.lexsort((pts[:, 2], pts[:, 1], pts[:, 0]))]
sd = np.concatenate([pts, np.arange(2, 18)[:, None]], 1).astype(np.uint32)
for mode in (4, 2):
    ctx.setFloodMode(mode)
    for lv in (4, 6, 8, 12, 16, 24, 32):
        ctx.setFloodLevels(lv)
        out = []
        for name, grid, s, df in (("vessel-manh", g, seeds, 1), ("vessel-cheb", g, seeds, 2), ("vessel-cheb-extra", g, xs, 2), ("dense-manh", g2, sd, 1)):
            fl = vf.FloodFracturer(); fl.setDistanceFunction(df)
            best = 1e9
            for rep in range(3):
                if grid is g: grid.updateSSBO(occ)
                else: grid.fillValue(1)
                ctx.synchronize(); ctx.timer_start(); fl.build(grid, s); best = min(best, ctx.timer_stop())
            out.append(f"{name} {best:.3f} ms ({fl.last_stats.tile_rounds} rounds, {fl.last_stats.tile_visits} visits)")
        print(f"ctas/sm {mode} levels {lv}: " + "; ".join(out), flush=True)
