"""CPU stand-in for one slab (tests only): same relax / boundary / ingest / finalize surface as voxelfragmentml_b200.slab.GpuSlab,
computed by the oracle.  Lets the slab protocol (partitioning, seed localisation, plane exchange, termination) run under gloo."""
import numpy as np
import torch

WALL, UNREACHED = 0xFFFFFFFF, 0xFFFFFFFE


class OracleSlab:
    def __init__(self, orc, labels_with_halo, seeds_global, x0, x1, X, dfunc):
        from voxelfragmentml_b200.slab import localise_seeds

        self.orc = orc
        self.nneigh = 6 if dfunc == 1 else 26
        self.seeds_global = np.asarray(seeds_global, np.uint32)
        self.keys = np.where(labels_with_halo == 0, WALL, UNREACHED).astype(np.uint32)
        for lx, y, z, order in localise_seeds(self.seeds_global, x0, x1):
            self.keys[lx, y, z] = order
        self.plane = self.keys.shape[1] * self.keys.shape[2]
        self.recv = [torch.empty(self.plane, dtype=torch.int32) for _ in range(2)]

    def relax(self):
        before = self.keys[1:-1].copy()
        self.orc.relax_keys_slab(self.keys, self.nneigh)
        return int((before != self.keys[1:-1]).sum())

    def boundary(self, side):
        k = self.keys[1] if side == 0 else self.keys[-2]
        return torch.from_numpy(k.reshape(-1).view(np.int32))

    def ingest(self, side, plane):
        recv = plane.numpy().view(np.uint32).reshape(self.keys.shape[1:])
        halo = self.keys[0] if side == 0 else self.keys[-1]
        m = recv < halo
        halo[m] = recv[m]
        return int(m.sum())

    def finalize(self):
        k = self.keys[1:-1]
        out = np.where(k == WALL, 0, 1).astype(np.uint16)
        reached = k < UNREACHED
        out[reached] = self.seeds_global[(k[reached] & 0x7FFF).astype(np.int64), 3].astype(np.uint16)
        return out
