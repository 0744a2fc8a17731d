"""Design aid: cfg4 batch with 16 jobs, pool of 8 shapes vs every mesh its own shape; per-job wall time of each mesh (first vs later meshes)."""
import os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import voxelfragmentml_b200 as vf
from voxelfragmentml_b200 import synth
lib = vf._capi.load()
jobs, meshes = 16, 128
for pool_mode in ("pool8", "distinct", "distinct-again"):
    shapes = {m: synth.vessel_mesh(m % 8 if pool_mode == "pool8" else m) for m in range(meshes)}
    if pool_mode != "distinct-again":
        workers = []
        for j in range(jobs):
            c = vf.Context(0); c.setFloodLevels(8); c.setFloodMode(0)
            g = vf.RegularGrid(c, (256, 256, 256)); c.reserve((256, 256, 256)); workers.append((c, g))
    times = [[] for _ in range(jobs)]
    def work(j):
        c, g = workers[j]
        for m in range(j, meshes, jobs):
            v, f = shapes[m]
            t0 = time.perf_counter()
            mn, mx = synth.mesh_aabb(v); dims = np.zeros(3, np.uint32); lib.vf_dims_rule(mn.ctypes.data, mx.ctypes.data, 256, dims.ctypes.data)
            g.setAABB(mn, mx, tuple(int(d) for d in dims)); g.fill(v, f); c.initSeed(80 + m)
            t1 = time.perf_counter()
            for k in range(10):
                nf = 2 + (k % 9)
                g.homogenize(); vf.fracture_model(g, vf.FractureParameters(_numSeeds=nf, _numExtraSeeds=2 * nf)); g.countValuesUndoMask()
            times[j].append(((t1 - t0) * 1e3, (time.perf_counter() - t1) * 1e3))
        c.synchronize()
    t0 = time.perf_counter()
    ts = [threading.Thread(target=work, args=(j,)) for j in range(jobs)]
    [t.start() for t in ts]; [t.join() for t in ts]
    dt = time.perf_counter() - t0
    vox = np.array([[a for a, b in tj] for tj in times]); fr = np.array([[b for a, b in tj] for tj in times])
    print(f"{pool_mode}: {meshes / dt:.1f} models/s; voxelize+setup ms by position in the job's list: {np.round(vox.mean(0), 1)}; 10 fragmentations ms: {np.round(fr.mean(0), 1)}", flush=True)
