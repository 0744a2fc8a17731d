"""Slab-partitioned flood fill over several GPUs (BASELINE config 5; new — the reference is single-GPU).

The grid is cut into slabs of the slowest axis x.  The flood's key field has a schedule-independent fixed point, so every rank
relaxes its slab to a local fixed point, the owned boundary planes travel to the neighbours (NCCL send/recv over NVLink through
torch.distributed, or plain copies when several slabs live in one process), the received planes are ingested, and the loop ends
when an all-reduced change counter is zero.  This module holds the host logic only; every cell is touched by libvoxfrag kernels
(or, in the CPU tests, by a stand-in backend with the same four methods)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from ._capi import check, ptr


def partition(X: int, world: int):
    """Balanced contiguous x-ranges [(x0, x1), ...], every rank gets at least one plane."""
    if world > X:
        raise ValueError("more slabs than planes")
    base, rem = divmod(X, world)
    out, x = [], 0
    for r in range(world):
        n = base + (1 if r < rem else 0)
        out.append((x, x + n))
        x += n
    return out


def localise_seeds(seeds_global, x0: int, x1: int):
    """Seeds whose x lies in the slab or its halo planes -> {local x (halo included), y, z, global order}."""
    s = np.asarray(seeds_global, dtype=np.int64)
    order = np.arange(len(s))
    keep = (s[:, 0] >= x0 - 1) & (s[:, 0] <= x1)
    loc = np.stack([s[keep, 0] - (x0 - 1), s[keep, 1], s[keep, 2], order[keep]], axis=1)
    return np.ascontiguousarray(loc, dtype=np.uint32)


def slab_with_halo(grid_global: np.ndarray, x0: int, x1: int) -> np.ndarray:
    """Planes x0-1 .. x1 of a host grid, EMPTY where the global grid ends (test / small-scale helper)."""
    X = grid_global.shape[0]
    out = np.zeros((x1 - x0 + 2,) + grid_global.shape[1:], dtype=grid_global.dtype)
    lo, hi = max(x0 - 1, 0), min(x1 + 1, X)
    out[lo - (x0 - 1) : hi - (x0 - 1)] = grid_global[lo:hi]
    return out


class GpuSlab:
    """One slab on one GPU.  `labels_with_halo`: host uint16 [(xs+2), Y, Z]; keys live in a torch tensor so NCCL can send planes."""

    def __init__(self, ctx, labels_with_halo, seeds_global, x0: int, x1: int, X: int, dfunc: int, shape=None, defer_init: bool = False):
        """labels_with_halo: host uint16 [(xs+2), Y, Z], or a callable fill(grid) that writes the slab on the device (then pass
        `shape` = ((xs+2), Y, Z))."""
        import torch

        from .api import RegularGrid

        self.ctx, self.x0, self.x1 = ctx, x0, x1
        self.seeds_global = np.ascontiguousarray(seeds_global, dtype=np.uint32)
        self.shape = tuple(shape) if callable(labels_with_halo) else tuple(labels_with_halo.shape)
        self.grid = RegularGrid(ctx, self.shape)
        if callable(labels_with_halo):
            labels_with_halo(self.grid)
        else:
            self.grid.updateSSBO(labels_with_halo)
        self.plane = self.shape[1] * self.shape[2]
        self.keys = torch.empty(self.shape[0] * self.plane, dtype=torch.int32, device=f"cuda:{ctx.device}")
        self.recv = [torch.empty(self.plane, dtype=torch.int32, device=self.keys.device) for _ in range(2)]
        self._init_args = (int(dfunc), int(x0 > 0), int(x1 < X))
        self._h = None
        if not defer_init:
            self.start()

    def start(self):
        """key-field initialisation + seed planting (FloodFracturer.cpp:99-132); separate so that benchmarks can time it"""
        loc = localise_seeds(self.seeds_global, self.x0, self.x1)
        h = C.c_void_p()
        dfunc, has_lo, has_hi = self._init_args
        check(self.ctx._lib.vf_flood_slab_init(self.grid._h, C.c_void_p(self.keys.data_ptr()), ptr(loc) if len(loc) else None, len(loc), dfunc, has_lo, has_hi,
                                               C.byref(h)))
        self._h = h

    def relax(self) -> int:
        n = C.c_uint64(0)
        check(self.ctx._lib.vf_flood_slab_relax(self._h, C.byref(n)))
        return int(n.value)

    def boundary(self, side: int):
        """owned plane next to the lo (0) / hi (1) halo, as a contiguous torch view of the key field"""
        k = self.keys.view(self.shape[0], self.plane)
        return k[1] if side == 0 else k[self.shape[0] - 2]

    def ingest(self, side: int, plane) -> int:
        n = C.c_uint64(0)
        check(self.ctx._lib.vf_flood_slab_ingest(self._h, side, C.c_void_p(plane.data_ptr()), C.byref(n)))
        return int(n.value)

    def run_native(self, comm, rank: int, world: int):
        """the whole exchange loop inside libvoxfrag (C++ over NCCL, vf_flood_slab_run); comm from nccl_comm() below, None when world == 1"""
        iters, moved = C.c_uint32(0), C.c_uint64(0)
        check(self.ctx._lib.vf_flood_slab_run(self._h, comm, rank, world, C.byref(iters), C.byref(moved)))
        return int(iters.value), int(moved.value)

    def finalize(self, download: bool = True):
        md = C.c_uint32(0)
        check(self.ctx._lib.vf_flood_slab_finalize(self._h, ptr(self.seeds_global), len(self.seeds_global), C.byref(md)))
        self.max_dist = md.value
        return self.grid.updateGrid()[1:-1] if download else None

    def close(self):
        if getattr(self, "_h", None):
            self.ctx._lib.vf_flood_slab_destroy(self._h)
            self._h = None
            self.grid.close()


class LabelSlab:
    """One slab of a label grid on one GPU for the operators that need no iteration across slabs: nearest-seed
    fragmentation (pointwise: the halo planes are computed), detectBoundaries / erosion passes / the 3^3 sweep (one-cell stencils: the
    halo planes are exchanged after every pass).  The slab lives in a torch tensor [(halo?) + xs + (halo?), Y, Z] so that its planes can
    travel (copies inside one process, NCCL send / recv through torch.distributed across ranks).  Connected-to-seed cleanup (C1) is
    a whole-grid connectivity question and is not sharded this way."""

    def __init__(self, ctx, occupancy_with_halo, x0: int, x1: int, X: int, shape=None):
        """occupancy_with_halo: host uint16 [planes, Y, Z] = planes x0 - (x0 > 0) .. x1 - 1 + (x1 < X) of the whole grid, or a callable
        fill(grid) that writes those planes on the device (then pass `shape` = (planes, Y, Z))"""
        import torch

        from .api import RegularGrid

        self.ctx, self.x0, self.x1, self.X = ctx, x0, x1, X
        self.has_lo, self.has_hi = int(x0 > 0), int(x1 < X)
        self.origin = x0 - self.has_lo  # plane of the whole grid the slab's plane 0 holds
        if callable(occupancy_with_halo):
            self.shape = tuple(int(d) for d in shape)
            self.t = torch.empty(self.shape, dtype=torch.int16, device=f"cuda:{ctx.device}")
        else:
            host = np.ascontiguousarray(occupancy_with_halo, np.uint16)
            self.shape = tuple(host.shape)
            self.t = torch.from_numpy(host.view(np.int16)).to(f"cuda:{ctx.device}")
        assert self.shape[0] == (x1 - x0) + self.has_lo + self.has_hi
        self.grid = RegularGrid(ctx, self.shape, device_ptr=self.t.data_ptr())
        if callable(occupancy_with_halo):
            torch.cuda.synchronize()
            occupancy_with_halo(self.grid)
            ctx.synchronize()
        torch.cuda.synchronize()  # the upload ran on torch's stream, the kernels run on the context's

    def naive(self, seeds_global, dfunc: int):
        s = np.ascontiguousarray(seeds_global, np.uint32)
        check(self.ctx._lib.vf_fracture_naive_slab(self.grid._h, ptr(s), len(s), int(dfunc), self.origin, self.X))

    def detect_boundaries(self, size: int = 1):
        self.grid.detectBoundaries(size)

    def erode_pass(self, etype, size, prob, thr, noise, boundary_mode=0):
        noise = np.ascontiguousarray(noise, np.float32)
        offset = self.origin * self.shape[1] * self.shape[2]
        check(self.ctx._lib.vf_erode_pass(self.grid._h, int(etype), int(size), float(prob), float(thr), ptr(noise), len(noise), int(boundary_mode), offset))

    def sweep(self):
        self.grid.removeIsolatedRegions()

    def boundary(self, side: int):
        """owned plane next to the lo (0) / hi (1) halo"""
        return self.t[1] if side == 0 else self.t[self.shape[0] - 2]

    def halo(self, side: int):
        return self.t[0] if side == 0 else self.t[self.shape[0] - 1]

    def owned(self):
        """host copy of the owned planes"""
        self.ctx.synchronize()
        a = self.t.cpu().numpy().view(np.uint16)
        return a[self.has_lo : a.shape[0] - self.has_hi]

    def close(self):
        self.grid.close()


def exchange_labels_local(slabs):
    """halo planes <- the neighbours' owned boundary planes, all slabs in one process"""
    import torch

    for s in slabs:
        s.ctx.synchronize()
    for r in range(len(slabs) - 1):
        slabs[r].halo(1).copy_(slabs[r + 1].boundary(0))
        slabs[r + 1].halo(0).copy_(slabs[r].boundary(1))
    torch.cuda.synchronize()


def exchange_labels(slab, rank: int, world: int, dist):
    """one slab per rank: the owned boundary planes travel with dist.batch_isend_irecv (NCCL over NVLink on a GPU box).  `slab` needs
    boundary(side) / halo(side) returning 16-bit planes and, when it computes on a stream of its own, ctx.synchronize()."""
    import torch

    if getattr(slab, "ctx", None) is not None:
        slab.ctx.synchronize()
    ops, cuda = [], False
    for side, peer in ((0, rank - 1), (1, rank + 1)):
        if 0 <= peer < world:
            # planes travel as bytes: NCCL has no 16-bit integer type
            ops.append(dist.P2POp(dist.isend, slab.boundary(side).contiguous().view(torch.uint8), peer))
            ops.append(dist.P2POp(dist.irecv, slab.halo(side).view(torch.uint8), peer))
            cuda = cuda or slab.halo(side).is_cuda
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        if cuda:
            torch.cuda.synchronize()


def erode_slabs(slabs, etype, size, iterations, prob, thr, noise, boundary_mode=0, exchange=None):
    """RegularGrid::erode (RegularGrid.cpp:82-159) over slabs: detectBoundaries once, an erosion pass per iteration, the 3^3 sweep, with a
    halo exchange after every pass.  `exchange()` defaults to copies between the slabs of this process."""
    ex = exchange or (lambda: exchange_labels_local(slabs))
    for s in slabs:
        s.detect_boundaries(1)
    ex()
    for _ in range(iterations):
        for s in slabs:
            s.erode_pass(etype, size, prob, thr, noise, boundary_mode)
        ex()
    for s in slabs:
        s.sweep()
    ex()


def nccl_comm(ctx, rank: int, world: int, dist):
    """An NCCL communicator owned by libvoxfrag for the native exchange loop: rank 0 draws the unique id (vf_nccl_unique_id), the 128 bytes
    travel through torch.distributed (any backend), every rank joins (vf_nccl_comm_create).  Free with ctx._lib.vf_nccl_comm_destroy."""
    import torch

    buf = (C.c_ubyte * 128)()
    if rank == 0:
        check(ctx._lib.vf_nccl_unique_id(buf))
    t = torch.tensor(list(buf), dtype=torch.uint8)
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.broadcast(t, 0)
    raw = bytes(t.cpu().tolist())
    comm = C.c_void_p()
    check(ctx._lib.vf_nccl_comm_create(ctx._h, raw, world, rank, C.byref(comm)))
    return comm


def run_local(slabs):
    """All slabs in one process (one GPU or a CPU stand-in): exchange = direct copies.  Returns (iterations, halo bytes moved)."""
    iters, moved = 0, 0
    while True:
        iters += 1
        changed = sum(s.relax() for s in slabs)
        for r in range(len(slabs) - 1):
            up, dn = slabs[r].boundary(1), slabs[r + 1].boundary(0)
            moved += 2 * up.numel() * 4
            changed += slabs[r + 1].ingest(0, up)
            changed += slabs[r].ingest(1, dn)
        if changed == 0:
            return iters, moved


def run_distributed(slab, rank: int, world: int, dist, make_recv=None):
    """One slab per rank; planes travel with dist.batch_isend_irecv, the change counter with all_reduce.
    `slab` needs relax / boundary / ingest; `make_recv(side)` returns a receive buffer (defaults to slab.recv[side])."""
    import torch

    iters, moved = 0, 0
    while True:
        iters += 1
        changed = slab.relax()
        ops, got = [], []
        for side, peer in ((0, rank - 1), (1, rank + 1)):
            if 0 <= peer < world:
                buf = slab.recv[side] if make_recv is None else make_recv(side)
                ops.append(dist.P2POp(dist.isend, slab.boundary(side), peer))
                ops.append(dist.P2POp(dist.irecv, buf, peer))
                got.append((side, buf))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
            if got and got[0][1].is_cuda:
                torch.cuda.synchronize()
        for side, buf in got:
            moved += buf.numel() * 4
            changed += slab.ingest(side, buf)
        t = torch.tensor([changed], dtype=torch.int64, device=got[0][1].device if got else "cpu")
        dist.all_reduce(t)
        if int(t.item()) == 0:
            return iters, moved
