"""torchrun target of tests/test_slab_gpu.py::test_label_slabs_over_nccl (one rank per GPU): nearest-seed fragmentation and RegularGrid::erode on
a grid cut into slabs along x, halo planes exchanged over NCCL (torch.distributed), every rank's slab against the oracle on the whole grid.
Prints SLAB_LABELS_OK on rank 0."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import oracle as orc
    import voxelfragmentml_b200 as vf
    from conftest import pick_seeds, random_blob_grid
    from voxelfragmentml_b200 import slab

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = vf.Context(local)
    for dims, dfunc in (((96, 64, 128), 0), ((75, 50, 60), 1)):
        g = random_blob_grid(dims, 5, fill=0.8).astype(np.uint16)  # seeded: the same grid on every rank
        seeds = pick_seeds(g, 24, 4)
        noise = orc.Rng(1080).fill_noise(50001)
        X = dims[0]
        x0, x1 = slab.partition(X, world)[rank]
        s = slab.LabelSlab(ctx, g[x0 - int(x0 > 0) : x1 + int(x1 < X)], x0, x1, X)
        s.naive(seeds, dfunc)
        want = orc.naive(g.copy(), seeds, dfunc)
        ok = np.array_equal(s.owned(), want[x0:x1])
        slab.erode_slabs([s], 1, 3, 3, 0.5, 0.5, noise, 0, exchange=lambda: slab.exchange_labels(s, rank, world, dist))
        wantE = orc.erode(want.copy(), noise, 1, 3, 3, 0.5, 0.5, boundary_mode=0)
        ok = ok and np.array_equal(s.owned(), wantE[x0:x1])
        s.close()
        t = torch.tensor([int(ok)], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        if int(t.item()) != 1:
            raise SystemExit(f"rank {rank}: label slab differs at dims={dims}")
        if rank == 0:
            print(f"dims={dims} dfunc={dfunc} world={world}: naive + erode(ELLIPSE,3,3it) per slab identical to the whole grid", flush=True)
    ctx.close()
    dist.destroy_process_group()
    if rank == 0:
        print("SLAB_LABELS_OK", flush=True)


if __name__ == "__main__":
    main()
