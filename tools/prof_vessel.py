"""Design aid: per-stage times of cfg3 on the SAT-voxelized synthetic vessel (352 x 512 x 352, 64 OUTER seeds)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import voxelfragmentml_b200 as vf
from voxelfragmentml_b200 import synth
from bench import noise_table
ctx = vf.Context(0)
res = int(sys.argv[1]) if len(sys.argv) > 1 else 512
v, f = synth.vessel_mesh(0); mn, mx = synth.mesh_aabb(v)
dims = np.zeros(3, np.uint32); vf._capi.load().vf_dims_rule(mn.ctypes.data, mx.ctypes.data, res, dims.ctypes.data); dims = tuple(int(d) for d in dims)
N = int(np.prod(dims))
work = torch.empty(N, dtype=torch.int16, device="cuda")
g = vf.RegularGrid(ctx, dims, device_ptr=work.data_ptr()); g.setAABB(mn, mx, dims); g.fill(v, f)
ctx.initSeed(80); seeds = vf.Seeder.uniform(g, 64)
pristine = work.clone()
noise = noise_table(1080, 1000000)
nv = vf.NaiveFracturer(); nv.setDistanceFunction(0)
if len(sys.argv) > 2: ctx.setC1Mode(int(sys.argv[2]))
for rep in range(3):
    work.copy_(pristine); torch.cuda.synchronize()
    for name, fn in (("naive", lambda: nv.build(g, seeds)), ("c1", lambda: vf.NaiveFracturer.removeIsolatedRegions(g, seeds)),
                     ("erode", lambda: g.erode(1, 3, 3, 0.5, 0.5, noise=noise)), ("hist", lambda: g.countValuesUndoMask())):
        ctx.timer_start(); fn(); ms = ctx.timer_stop()
        print(f"rep {rep} {name}: {ms:.3f} ms", flush=True)
