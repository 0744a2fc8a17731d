"""Design aid (CPU): the "descent certificate" formulation of C1 (NaiveFracturer::removeIsolatedRegionsCPU semantics) checked against the
oracle.  A labelled cell that has a same-label 6-neighbour one Manhattan step closer to its own seed is connected to the seed if that
neighbour is; so only the cells WITHOUT such a neighbour (F), the cells all of whose descent neighbours are dead ends (closure D of F), and
the same-label cells around them need a connectivity search — a few dozen cells on the cfg3 grid instead of a union-find over 134 M.
    kept = cells outside D  +  cells of D reachable inside D from a D-cell that touches a same-label cell outside D
usage: python tools/c1_descent_prototype.py [n=128]   (cfg3-dense labels at n^3, the golden vessel, porous blobs)"""
import os, sys
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

NB = [(1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1)]


def c1_descent(grid, seeds):
    """returns (result grid, |F|, |D|).  Vectorised certificate pass, Python sets for the (small) rest."""
    g = grid.copy()
    X, Y, Z = g.shape
    for x, y, z, w in seeds:  # NaiveFracturer.cpp:120-123: the seed cell carries the seed's label, a later seed wins
        g[x, y, z] = w
    pos = {}  # every seed is the start of its own label, also one whose cell a later seed took (the reference's front holds all seeds)
    for x, y, z, w in seeds:
        assert int(w) not in pos, "two seeds with one label: the CUDA path falls back to the union-find"
        pos[int(w)] = (int(x), int(y), int(z))
    table = np.full((65536, 3), -(10 ** 6), np.int64)  # labels without a seed: no descent neighbour can exist
    has_seed = np.zeros(65536, bool)
    for w, p in pos.items():
        table[w] = p
        has_seed[w] = True
    active = g > 1
    P = table[g]  # (X, Y, Z, 3)
    C = np.stack(np.meshgrid(np.arange(X), np.arange(Y), np.arange(Z), indexing="ij"), axis=-1)
    step = np.sign(C - P)
    cert = np.zeros(g.shape, bool)
    idx = [C[..., 0], C[..., 1], C[..., 2]]
    for a in range(3):
        nidx = list(idx)
        nidx[a] = np.clip(idx[a] - step[..., a], 0, g.shape[a] - 1)
        cert |= (step[..., a] != 0) & (g[nidx[0], nidx[1], nidx[2]] == g)
    cert |= np.abs(C - P).sum(axis=-1) == 1  # next to the start: entered from it whatever label the start's cell carries
    cert &= has_seed[g]                      # a label without a seed has no start at all
    is_seed = (step == 0).all(axis=-1) & has_seed[g]
    F = active & ~cert & ~is_seed
    D = set(map(tuple, np.argwhere(F)))
    nF = len(D)

    def man(c, p):
        return abs(c[0] - p[0]) + abs(c[1] - p[1]) + abs(c[2] - p[2])

    def same_label_neighbours(c):
        for d in NB:
            n = (c[0] + d[0], c[1] + d[1], c[2] + d[2])
            if 0 <= n[0] < X and 0 <= n[1] < Y and 0 <= n[2] < Z and g[n] == g[c]:
                yield n

    frontier = list(D)
    while frontier:  # closure: a cell all of whose descent neighbours are in D is in D
        nxt = []
        for u in frontier:
            L = int(g[u])
            if L not in pos:
                continue
            for v in same_label_neighbours(u):
                if v in D or v == pos[L] or man(v, pos[L]) != man(u, pos[L]) + 1:
                    continue
                desc = [n for n in same_label_neighbours(v) if man(n, pos[L]) == man(v, pos[L]) - 1]
                if all(n in D for n in desc):
                    D.add(v)
                    nxt.append(v)
        frontier = nxt
    alive = {u for u in D if any(n not in D for n in same_label_neighbours(u))}
    frontier = list(alive)
    while frontier:
        nxt = []
        for u in frontier:
            for n in same_label_neighbours(u):
                if n in D and n not in alive:
                    alive.add(n)
                    nxt.append(n)
        frontier = nxt
    out = g.copy()
    out[~active] = 0  # FREE cells are dropped (the reference rebuilds from an all-EMPTY grid)
    for u in D - alive:
        out[u] = 0
    return out, nF, len(D)


if __name__ == "__main__":
    import bench
    import oracle as orc
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
    from conftest import random_blob_grid, pick_seeds

    orc.use_all_cores()
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    cases = []
    seeds = bench.synth_seeds_dense(n, 64, bench.rng_uniform_stream(80))
    for df in (0, 1, 2):
        cases.append((f"dense {n}^3 naive dfunc {df}", orc.naive(np.ones((n, n, n), np.uint16), seeds, df), seeds))
    occ = orc.decode_rle(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "AL_12B_grid_128r.rle"), "rb").read())
    vs, _ = orc.seed_uniform(orc.Rng(80), occ, 8)
    cases.append(("golden vessel, 8 seeds", orc.naive(occ.copy(), vs, 0), vs))
    for k in range(3):
        b = random_blob_grid((40, 36, 44), k, fill=0.55, smooth=1)
        bs = pick_seeds(b, 6, k)
        cases.append((f"porous blob {k}", orc.naive(b.copy(), bs, k % 3), bs))
    for name, lab, sd in cases:
        want = orc.remove_isolated_regions_cpu(lab.copy(), sd)
        got, nF, nD = c1_descent(lab, sd)
        print(f"{name}: |F| = {nF}, |D| = {nD}, removed = {int((want != lab).sum())}, identical = {np.array_equal(got, want)}")
        assert np.array_equal(got, want)
