"""Profiling helper (not part of the product): runs selected cfg3 stages once on a dense n^3 grid so that `ncu` can list the
kernels of one stage.  usage: python tools/prof_stage.py 512 naive,c1,erode,hist"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import voxelfragmentml_b200 as vf
from bench import synth_seeds_dense, rng_uniform_stream, noise_table

n = int(sys.argv[1]); stages = sys.argv[2].split(",")
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
ctx = vf.Context(0)
work = torch.ones(n**3, dtype=torch.int16, device="cuda")
grid = vf.RegularGrid(ctx, (n, n, n), device_ptr=work.data_ptr())
ctx.reserve((n, n, n))
seeds = synth_seeds_dense(n, 64, rng_uniform_stream(80))
noise = noise_table(1080, 1000000)
nv = vf.NaiveFracturer(); nv.setDistanceFunction(0)
fl = vf.FloodFracturer(); fl.setDistanceFunction(1)
for rep in range(reps):
    work.fill_(1); torch.cuda.synchronize()
    for s in stages:
        ctx.timer_start()
        if s == "naive": nv.build(grid, seeds)
        elif s == "c1": vf.NaiveFracturer.removeIsolatedRegions(grid, seeds)
        elif s == "erode": grid.erode(1, 3, 3, 0.5, 0.5, noise=noise)
        elif s == "hist": grid.countValues()
        elif s == "flood": fl.build(grid, seeds[:16])
        elif s == "flood26": fl.setDistanceFunction(2); fl.build(grid, seeds[:16])
        elif s == "detect": grid.detectBoundaries(1)
        ms = ctx.timer_stop()
        print(f"rep {rep} {s}: {ms:.3f} ms", flush=True)
    if "flood" in stages or "flood26" in stages:
        st = fl.last_stats; print("flood stats", st.tile_rounds, st.tile_visits, st.max_dist)
