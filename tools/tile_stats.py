"""Design aid (CPU only, uses the oracle as a label generator): how much of the cfg3-dense label grid is uniform at the granularities the
stencil / component-labelling kernels skip work at.  usage: python tools/tile_stats.py [n=512]
Printed for the labels after the naive stage: fraction of 8x8x64 stencil tiles whose staged region (tile + 1-cell halo) holds one value
(the kernels' whole-tile shortcut), fraction of tiles uniform without the halo, fraction of 8-cell chunks with a uniform 3x3x10 window
(the chunks that are copied instead of evaluated), and the same for 8x8x8 bricks with an 8-cell halo (a candidate "inert for the whole
erode call" granularity)."""
import os, sys, time
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import oracle as orc

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
orc.use_all_cores()
seeds = bench.synth_seeds_dense(n, bench.CFG3["nseeds"], bench.rng_uniform_stream(bench.CFG3["rng_seed"]))
g = orc.naive(np.ones((n, n, n), np.uint16), seeds, orc.EUCLIDEAN)


def box_uniform(g, bx, by, bz, halo):
    """per (bx, by, bz) block: does block + halo (clipped at the grid faces by padding with 0) hold one value?"""
    p = np.zeros(tuple(s + 2 * halo for s in g.shape), np.uint16)
    p[halo:-halo, halo:-halo, halo:-halo] = g
    X, Y, Z = g.shape
    mn = np.full((X // bx, Y // by, Z // bz), 65535, np.uint16)
    mx = np.zeros_like(mn)
    # min / max over z windows first (the long axis), then over the x and y offsets
    zmn = np.stack([p[:, :, k * bz:k * bz + bz + 2 * halo].min(axis=2) for k in range(Z // bz)], axis=2)
    zmx = np.stack([p[:, :, k * bz:k * bz + bz + 2 * halo].max(axis=2) for k in range(Z // bz)], axis=2)
    for dx in range(bx + 2 * halo):
        for dy in range(by + 2 * halo):
            mn = np.minimum(mn, zmn[dx:dx + X - bx + 1:bx, dy:dy + Y - by + 1:by])
            mx = np.maximum(mx, zmx[dx:dx + X - bx + 1:bx, dy:dy + Y - by + 1:by])
    return mn == mx


t0 = time.time()
print(f"{n}^3 dense, {len(seeds)} seeds, labels after the naive stage")
print("  8x8x64 tiles, staged region (halo 1) uniform : %.3f" % box_uniform(g, 8, 8, 64, 1).mean())
t = g.reshape(n // 8, 8, n // 8, 8, n // 64, 64)
print("  8x8x64 tiles uniform without halo            : %.3f" % (t.min(axis=(1, 3, 5)) == t.max(axis=(1, 3, 5))).mean())
print("  16x16x32 tiles uniform without halo (C1)     : %.3f" % (lambda u: (u.min(axis=(1, 3, 5)) == u.max(axis=(1, 3, 5))).mean())(g.reshape(n // 16, 16, n // 16, 16, n // 32, 32)))
print("  8-cell chunks with a uniform 3x3x10 window   : %.3f" % box_uniform(g, 1, 1, 8, 1).mean())
print("  8x8x8 bricks uniform with an 8-cell halo     : %.3f" % box_uniform(g, 8, 8, 8, 8).mean())
print("  8x8x8 bricks uniform with a 4-cell halo      : %.3f" % box_uniform(g, 8, 8, 8, 4).mean())
print("  (%.0f s)" % (time.time() - t0))
print("  8x8x64 tiles uniform with an 8-cell halo      : %.3f" % box_uniform(g, 8, 8, 64, 8).mean())
print("  8x8x32 tiles uniform with an 8-cell halo      : %.3f" % box_uniform(g, 8, 8, 32, 8).mean())
print("  8x8x16 tiles uniform with an 8-cell halo      : %.3f" % box_uniform(g, 8, 8, 16, 8).mean())
