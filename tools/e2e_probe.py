"""Timeline probe for bench.py's pipelined e2e loop: CUDA events around each step's upload / kernels / download, per slot stream."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import voxelfragmentml_b200 as vf
from bench import synth_seeds_dense, rng_uniform_stream, noise_table, CFG3

n = 512; N = n ** 3; dims = (n, n, n)
seeds = synth_seeds_dense(n, CFG3["nseeds"], rng_uniform_stream(80))
noise = torch.from_numpy(noise_table(1080, CFG3["nnoise"])).pin_memory().numpy()
h_in = torch.ones(N, dtype=torch.int16).pin_memory()
slots = []
for k in range(3):
    c = vf.Context(0); g = vf.RegularGrid(c, dims); c.reserve(dims)
    slots.append((c, g, torch.empty(N, dtype=torch.int16).pin_memory(), torch.cuda.ExternalStream(c.stream)))
naive = vf.NaiveFracturer(); naive.setDistanceFunction(CFG3["dfunc"])
et, es, ei, ep, eth = CFG3["erosion"]
ev = lambda: torch.cuda.Event(enable_timing=True)

def run(nsteps, log):
    base = ev(); base.record(slots[0][3])
    host0 = time.perf_counter()
    marks = []
    slots[0][1].upload_async(h_in)
    e = ev(); e.record(slots[0][3]); marks.append(("up_end", 0, e))
    for i in range(nsteps):
        c, g, ho, st = slots[i % 3]
        h = [time.perf_counter()]
        e = ev(); e.record(st); marks.append(("k_begin", i, e))
        naive.build(g, seeds); h.append(time.perf_counter())
        vf.NaiveFracturer.removeIsolatedRegions(g, seeds); h.append(time.perf_counter())
        g.erode(et, es, ei, ep, eth, noise=noise); h.append(time.perf_counter())
        if i + 1 < nsteps:
            c2, g2, _, st2 = slots[(i + 1) % 3]
            e = ev(); e.record(st2); marks.append(("up_begin", i + 1, e))
            g2.upload_async(h_in)
            e = ev(); e.record(st2); marks.append(("up_end", i + 1, e))
        h.append(time.perf_counter())
        g.countValues(); h.append(time.perf_counter())
        g.undoMask()
        e = ev(); e.record(st); marks.append(("k_end", i, e))
        g.download_async(ho)
        e = ev(); e.record(st); marks.append(("down_end", i, e))
        h.append(time.perf_counter())
        if log:
            print(f"host step {i}: start {1e3*(h[0]-host0):7.2f} naive {1e3*(h[1]-h[0]):5.2f} c1 {1e3*(h[2]-h[1]):5.2f} erode {1e3*(h[3]-h[2]):5.2f} "
                  f"upload {1e3*(h[4]-h[3]):5.2f} hist {1e3*(h[5]-h[4]):5.2f} rest {1e3*(h[6]-h[5]):5.2f}")
    for c, *_ in slots:
        c.synchronize()
    torch.cuda.synchronize()
    if log:
        for name, i, e in marks:
            print(f"gpu  step {i}: {name:8s} at {base.elapsed_time(e):8.2f} ms")
    return time.perf_counter() - host0

run(3, False)
t = run(8, True)
print("per step", t / 8 * 1e3, "ms")
