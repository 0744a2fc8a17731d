"""torchrun target of tests/test_slab_gpu.py::test_native_exchange_loop_over_nccl (one rank per GPU): the C++ exchange loop over NCCL inside
libvoxfrag (vf_flood_slab_run) on the analytic solid vessel, every rank's slab against the single-context flood computed on rank 0's GPU,
which at 128^3 is itself checked against the oracle.  Prints SLAB_NATIVE_OK on rank 0."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import voxelfragmentml_b200 as vf
    from voxelfragmentml_b200 import slab, synth

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = vf.Context(local)
    lib = vf._capi.load()
    comm = slab.nccl_comm(ctx, rank, world, dist)
    for n, nseeds, dfunc in ((128, 256, 1), (256, 64, 2), (int(os.environ.get("VF_SLAB_TEST_N", "384")), 256, 1)):
        params = synth.solid_vessel_params(0)
        seeds = synth.solid_vessel_seeds(n, nseeds, params, 80)
        x0, x1 = slab.partition(n, world)[rank]

        def fill(grid, x0=x0, n=n, params=params):
            vf._capi.check(lib.vf_synth_solid_vessel(grid._h, x0 - 1, n, *params))

        s = slab.GpuSlab(ctx, fill, seeds, x0, x1, n, dfunc, shape=(x1 - x0 + 2, n, n))
        iters, moved = s.run_native(comm, rank, world)
        mine = s.finalize()
        s.close()
        # the whole grid in one context (every rank computes it: no gather of 2 B x n^3 needed)
        whole = vf.RegularGrid(ctx, (n, n, n))
        vf._capi.check(lib.vf_synth_solid_vessel(whole._h, 0, n, *params))
        occ = whole.updateGrid() if n == 128 else None
        fl = vf.FloodFracturer()
        fl.setDistanceFunction(dfunc)
        fl.build(whole, seeds, id_bits=15)
        single = whole.updateGrid()
        whole.close()
        ok = np.array_equal(mine, single[x0:x1]) and iters >= 2
        if n == 128 and rank == 0:
            import oracle as orc

            want, _ = orc.flood(occ.copy(), seeds, dfunc, id_bits=15)
            ok = ok and np.array_equal(single, want)
        t = torch.tensor([int(ok)], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        if int(t.item()) != 1:
            raise SystemExit(f"rank {rank}: slab labels differ at n={n}")
        if rank == 0:
            print(f"n={n} dfunc={dfunc} world={world}: {iters} exchange iterations, {moved} halo bytes per rank, labels identical", flush=True)
    lib.vf_nccl_comm_destroy(comm)
    ctx.close()
    dist.destroy_process_group()
    if rank == 0:
        print("SLAB_NATIVE_OK", flush=True)


if __name__ == "__main__":
    main()
