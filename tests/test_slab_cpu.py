"""Host logic of the multi-GPU slab flood on CPU: partition, seed localisation and the exchange protocol, exercised with a
gloo process group (world size 2 and 3) and an oracle-backed slab stand-in, against the single-address-space oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, pick_seeds, random_blob_grid


def test_partition_and_localise():
    from voxelfragmentml_b200 import slab

    assert slab.partition(10, 3) == [(0, 4), (4, 7), (7, 10)]
    assert slab.partition(2048, 8)[3] == (768, 1024)
    with pytest.raises(ValueError):
        slab.partition(2, 3)
    seeds = np.array([[0, 1, 2, 7], [3, 1, 1, 8], [4, 0, 0, 9], [9, 9, 9, 10]], np.uint32)
    loc = slab.localise_seeds(seeds, 4, 7)  # owns planes 4..6, halo planes 3 and 7
    assert loc.tolist() == [[0, 1, 1, 1], [1, 0, 0, 2]]
    g = np.arange(5 * 2 * 2, dtype=np.uint16).reshape(5, 2, 2)
    s = slab.slab_with_halo(g, 0, 2)
    assert s.shape == (4, 2, 2) and not s[0].any() and np.array_equal(s[1:], g[0:3])
    s = slab.slab_with_halo(g, 3, 5)
    assert not s[-1].any() and np.array_equal(s[:-1], g[2:5])


def _grid_and_seeds(orc):
    g = random_blob_grid((30, 14, 20), 21, fill=0.52, smooth=1)
    g[14:16, :, :10] = 0  # a partial wall right at a slab border forces detours through the neighbour slab
    return g, pick_seeds(g, 7, 4)


@pytest.mark.parametrize("dfunc", [1, 2])
@pytest.mark.parametrize("nslabs", [2, 4])
def test_local_protocol_matches_single_address_space(orc, dfunc, nslabs):
    from slab_backends import OracleSlab
    from voxelfragmentml_b200 import slab

    g, seeds = _grid_and_seeds(orc)
    want, _ = orc.flood(g.copy(), seeds, dfunc, id_bits=15)
    parts = slab.partition(g.shape[0], nslabs)
    slabs = [OracleSlab(orc, slab.slab_with_halo(g, x0, x1), seeds, x0, x1, g.shape[0], dfunc) for x0, x1 in parts]
    iters, moved = slab.run_local(slabs)
    got = np.concatenate([s.finalize() for s in slabs])
    assert np.array_equal(got, want)
    assert iters >= 2 and moved > 0


def _worker(rank, world, port, dfunc, outdir):
    import sys

    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle as orc
    from slab_backends import OracleSlab
    from voxelfragmentml_b200 import slab

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g, seeds = _grid_and_seeds(orc)
    x0, x1 = slab.partition(g.shape[0], world)[rank]
    s = OracleSlab(orc, slab.slab_with_halo(g, x0, x1), seeds, x0, x1, g.shape[0], dfunc)
    iters, moved = slab.run_distributed(s, rank, world, dist)
    np.save(os.path.join(outdir, f"labels_{rank}.npy"), s.finalize())
    np.save(os.path.join(outdir, f"meta_{rank}.npy"), np.array([iters, moved]))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,dfunc", [(2, 1), (3, 2)])
def test_gloo_protocol_matches_single_address_space(orc, tmp_path, world, dfunc):
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    mp.spawn(_worker, args=(world, port, dfunc, str(tmp_path)), nprocs=world, join=True)
    g, seeds = _grid_and_seeds(orc)
    want, _ = orc.flood(g.copy(), seeds, dfunc, id_bits=15)
    got = np.concatenate([np.load(tmp_path / f"labels_{r}.npy") for r in range(world)])
    assert np.array_equal(got, want)
    iters = [int(np.load(tmp_path / f"meta_{r}.npy")[0]) for r in range(world)]
    assert len(set(iters)) == 1 and iters[0] >= 2  # every rank leaves the loop in the same iteration


class _HostLabelSlab:
    """stand-in with LabelSlab's plane surface: planes x0 - (x0 > 0) .. x1 - 1 + (x1 < X) of a host grid as a torch tensor"""

    def __init__(self, g, x0, x1):
        X = g.shape[0]
        self.t = torch.from_numpy(g[x0 - int(x0 > 0) : x1 + int(x1 < X)].copy().view(np.int16))

    def boundary(self, side):
        return self.t[1] if side == 0 else self.t[self.t.shape[0] - 2]

    def halo(self, side):
        return self.t[0] if side == 0 else self.t[self.t.shape[0] - 1]


def _label_worker(rank, world, port, outdir):
    import sys

    sys.path.insert(0, ROOT)
    from voxelfragmentml_b200 import slab

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    X = 11
    g = (np.arange(X * 5 * 8, dtype=np.int64).reshape(X, 5, 8) % 60000).astype(np.uint16)
    x0, x1 = slab.partition(X, world)[rank]
    s = _HostLabelSlab(g, x0, x1)
    for side, peer in ((0, rank - 1), (1, rank + 1)):
        if 0 <= peer < world:
            s.halo(side).fill_(-1)
    slab.exchange_labels(s, rank, world, dist)
    np.save(os.path.join(outdir, f"slab_{rank}.npy"), s.t.numpy().view(np.uint16))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_label_halo_exchange(tmp_path, world):
    """exchange_labels (the halo exchange of the one-cell-halo operators, slab.LabelSlab): after it every halo plane holds the neighbour's
    owned boundary plane, i.e. every rank's tensor is again the slice of the whole grid it stands for."""
    from voxelfragmentml_b200 import slab

    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    mp.spawn(_label_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    X = 11
    g = (np.arange(X * 5 * 8, dtype=np.int64).reshape(X, 5, 8) % 60000).astype(np.uint16)
    for r, (x0, x1) in enumerate(slab.partition(X, world)):
        got = np.load(tmp_path / f"slab_{r}.npy")
        assert np.array_equal(got, g[x0 - int(x0 > 0) : x1 + int(x1 < X)])
