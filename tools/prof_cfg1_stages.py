"""Design aid: the stages of BASELINE cfg1 (128-max vessel, NAIVE EUCLIDEAN, 8 OUTER seeds) one by one, C1 in both modes."""
import sys
import numpy as np
sys.path.insert(0, ".")
import voxelfragmentml_b200 as vf
from voxelfragmentml_b200 import synth

ctx = vf.Context(0)
v, f = synth.vessel_mesh(0)
mn, mx = synth.mesh_aabb(v)
d = np.zeros(3, np.uint32)
ctx._lib.vf_dims_rule(vf._capi.ptr(np.float32(mn)), vf._capi.ptr(np.float32(mx)), int(sys.argv[1]) if len(sys.argv) > 1 else 128, vf._capi.ptr(d))
dims = tuple(int(x) for x in d)
g = vf.RegularGrid(ctx, dims)
g.setAABB(mn, mx, dims)
g.fill(v, f)
occ = g.updateGrid()
ctx.initSeed(80)
for rep in range(3):
    ctx.timer_start(); seeds = vf.Seeder.uniform(g, 8); t0 = ctx.timer_stop()
print("dims", dims, "occupied", int((occ != 0).sum()), f"seeding {t0:.3f} ms", flush=True)
nf = vf.NaiveFracturer()
nf.setDistanceFunction(0)
for mode in (0, 1):
    ctx.setC1Mode(mode)
    for rep in range(4):
        g.updateSSBO(occ); ctx.synchronize()
        ctx.timer_start(); nf.build(g, seeds); t1 = ctx.timer_stop()
        ctx.timer_start(); vf.NaiveFracturer.removeIsolatedRegions(g, seeds); t2 = ctx.timer_stop()
        ctx.timer_start(); g.detectBoundaries(1); t3 = ctx.timer_stop()
    print(f"C1 mode {mode}: naive {t1:.3f} ms, connected-to-seed {t2:.3f} ms, detectBoundaries {t3:.3f} ms", flush=True)
