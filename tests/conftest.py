import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def orc():
    import oracle

    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def golden_rle_bytes():
    return open(os.path.join(GOLDEN, "AL_12B_grid_128r.rle"), "rb").read()


@pytest.fixture(scope="session")
def vessel_grid(orc, golden_rle_bytes):
    """The reference's real 128x110x128 solid vessel occupancy (values {0,1})."""
    return orc.decode_rle(golden_rle_bytes)


def random_blob_grid(shape, seed, fill=0.55, smooth=2):
    """Deterministic random occupancy with caves and thin bridges (numpy only)."""
    rng = np.random.RandomState(seed)
    g = rng.rand(*shape) < fill
    for _ in range(smooth):
        p = np.pad(g, 1).astype(np.int32)
        s = sum(
            p[1 + dx : 1 + dx + shape[0], 1 + dy : 1 + dy + shape[1], 1 + dz : 1 + dz + shape[2]]
            for dx in (-1, 0, 1)
            for dy in (-1, 0, 1)
            for dz in (-1, 0, 1)
        )
        g = s >= 14
    return g.astype(np.uint16)


def pick_seeds(grid, n, seed):
    """n distinct occupied cells, sorted lexicographically, labels 2.. (same shape as Seeder::uniform output)."""
    rng = np.random.RandomState(seed)
    occ = np.argwhere(grid != 0)
    sel = occ[rng.choice(len(occ), size=n, replace=False)]
    sel = sel[np.lexsort((sel[:, 2], sel[:, 1], sel[:, 0]))]
    return np.concatenate([sel, np.arange(2, 2 + n)[:, None]], axis=1).astype(np.uint32)
