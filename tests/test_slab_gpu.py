"""cfg5 in miniature: slab-partitioned flood with the CUDA slab kernels (several slabs on one GPU, exchange by device copies)
against the single-address-space oracle, bit-exact."""
import numpy as np
import pytest

from conftest import pick_seeds, random_blob_grid

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import voxelfragmentml_b200 as vf

    c = vf.Context(0)
    yield c
    c.close()


def _run(ctx, g, seeds, dfunc, nslabs):
    from voxelfragmentml_b200 import slab

    import voxelfragmentml_b200 as vf

    parts = slab.partition(g.shape[0], nslabs)
    # a slab session owns its context's tile scratch, so every slab gets its own context (as it would on its own GPU)
    ctxs = [vf.Context(0) for _ in parts]
    slabs = [slab.GpuSlab(c, slab.slab_with_halo(g, x0, x1), seeds, x0, x1, g.shape[0], dfunc) for c, (x0, x1) in zip(ctxs, parts)]
    iters, moved = slab.run_local(slabs)
    got = np.concatenate([s.finalize() for s in slabs])
    for s in slabs:
        s.close()
    for c in ctxs:
        c.close()
    return got, iters, moved


@pytest.mark.parametrize("dfunc", [1, 2])
@pytest.mark.parametrize("nslabs", [2, 3, 8])
def test_vessel_slabs(ctx, orc, vessel_grid, dfunc, nslabs):
    seeds, _ = orc.seed_uniform(orc.Rng(80), vessel_grid, 16)
    want, st = orc.flood(vessel_grid.copy(), seeds, dfunc, id_bits=15)
    got, iters, moved = _run(ctx, vessel_grid, seeds, dfunc, nslabs)
    assert np.array_equal(got, want)
    assert iters >= 2


def test_many_labels_and_detours(ctx, orc):
    g = random_blob_grid((70, 40, 48), 31, fill=0.5, smooth=1)
    g[34:36, :, :30] = 0  # wall at the slab border: shortest paths leave the slab and come back
    seeds = pick_seeds(g, 300, 2)  # > 254 labels: 15-bit ids (cfg5 uses 256 seeds)
    for dfunc in (1, 2):
        want, _ = orc.flood(g.copy(), seeds, dfunc, id_bits=15)
        got, iters, _ = _run(ctx, g, seeds, dfunc, 2)
        assert np.array_equal(got, want)


def test_solid_vessel_256_seeds(ctx, orc):
    """the cfg5 workload at 1/16 linear scale: analytic solid vessel, Manhattan, 256 seeds, 8 slabs"""
    from voxelfragmentml_b200 import synth

    g = synth.solid_vessel_grid(128)
    seeds, _ = orc.seed_uniform(orc.Rng(80), g, 256, location=orc.BOTH)
    want, _ = orc.flood(g.copy(), seeds, 1, id_bits=15)
    got, iters, moved = _run(ctx, g, seeds, 1, 8)
    assert np.array_equal(got, want) and got.max() == 257


def test_native_loop_on_one_rank_equals_the_single_context_flood(ctx, orc, vessel_grid):
    """vf_flood_slab_run with world == 1 (no communicator): one slab = the whole grid"""
    from voxelfragmentml_b200 import slab

    seeds, _ = orc.seed_uniform(orc.Rng(80), vessel_grid, 16)
    for dfunc in (1, 2):
        want, _ = orc.flood(vessel_grid.copy(), seeds, dfunc, id_bits=15)
        X = vessel_grid.shape[0]
        s = slab.GpuSlab(ctx, slab.slab_with_halo(vessel_grid, 0, X), seeds, 0, X, X, dfunc)
        iters, moved = s.run_native(None, 0, 1)
        assert np.array_equal(s.finalize(), want) and iters >= 1 and moved == 0
        s.close()


def test_native_exchange_loop_over_nccl():
    """the C++ exchange loop over NCCL, one rank per GPU (needs >= 2 GPUs: `gpurun --gpus 2`); tests/slab_native_worker.py does the checking"""
    import os
    import subprocess
    import sys

    import torch

    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs at least two GPUs")
    world = 8 if ngpu >= 8 else (4 if ngpu >= 4 else 2)
    here = os.path.dirname(os.path.abspath(__file__))
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1", "--master-port", "29641",
                        os.path.join(here, "slab_native_worker.py")], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0 and "SLAB_NATIVE_OK" in p.stdout, p.stdout[-3000:] + p.stderr[-3000:]


def _label_slabs(g, nslabs):
    from voxelfragmentml_b200 import slab

    import voxelfragmentml_b200 as vf

    X = g.shape[0]
    parts = slab.partition(X, nslabs)
    ctxs = [vf.Context(0) for _ in parts]
    slabs = [slab.LabelSlab(c, g[x0 - int(x0 > 0) : x1 + int(x1 < X)], x0, x1, X) for c, (x0, x1) in zip(ctxs, parts)]
    return ctxs, slabs


@pytest.mark.parametrize("nslabs", [2, 3, 5])
@pytest.mark.parametrize("dims", [(50, 40, 64), (37, 26, 44), (33, 21, 27)])
def test_label_slabs_naive_and_erode_match_the_whole_grid(orc, nslabs, dims):
    """the operators that shard with a one-cell halo and no iteration.  Nearest-seed fragmentation per slab (halo planes computed,
    not exchanged) equals the oracle on the whole grid; detectBoundaries + erosion passes + the 3^3 sweep with a halo exchange after every
    pass equal RegularGrid::erode on the whole grid, noise indexed by the cell's position in the whole grid.  Tiled (Z % 8, Z % 4) and generic paths."""
    from voxelfragmentml_b200 import slab

    g = random_blob_grid(dims, 5, fill=0.8).astype(np.uint16)
    seeds = pick_seeds(g, 9, 4)
    noise = orc.Rng(1080).fill_noise(7001)
    for dfunc, bmode in ((0, 0), (1, 1), (2, 0)):
        want = orc.naive(g.copy(), seeds, dfunc)
        for k, (etype, iters, prob, thr) in enumerate(((1, 3, 0.5, 0.5), (0, 2, 0.9, 0.8))):
            ctxs, slabs = _label_slabs(g, nslabs)
            for s in slabs:
                s.naive(seeds, dfunc)
            if k == 0:
                got = np.concatenate([s.owned() for s in slabs])
                assert np.array_equal(got, want), f"naive dfunc {dfunc}: {int((got != want).sum())} cells differ"
                for r in range(len(slabs) - 1):  # the computed halo planes equal the neighbour's owned planes
                    assert np.array_equal(slabs[r].halo(1).cpu().numpy(), slabs[r + 1].boundary(0).cpu().numpy())
                    assert np.array_equal(slabs[r + 1].halo(0).cpu().numpy(), slabs[r].boundary(1).cpu().numpy())
            wantE = orc.erode(want.copy(), noise, etype, 3, iters, prob, thr, boundary_mode=bmode)
            slab.erode_slabs(slabs, etype, 3, iters, prob, thr, noise, bmode)
            gotE = np.concatenate([s.owned() for s in slabs])
            assert np.array_equal(gotE, wantE), f"erode type {etype} dfunc {dfunc} mode {bmode}: {int((gotE != wantE).sum())} cells differ"
            for s in slabs:
                s.close()
            for c in ctxs:
                c.close()


def test_label_slabs_over_nccl():
    """the same operators with one slab per GPU and the halo planes travelling over NCCL (needs >= 2 GPUs); tests/slab_labels_worker.py checks"""
    import os
    import subprocess
    import sys

    import torch

    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs at least two GPUs")
    world = 8 if ngpu >= 8 else (4 if ngpu >= 4 else 2)
    here = os.path.dirname(os.path.abspath(__file__))
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1", "--master-port", "29643",
                        os.path.join(here, "slab_labels_worker.py")], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0 and "SLAB_LABELS_OK" in p.stdout, p.stdout[-3000:] + p.stderr[-3000:]
