"""Deterministic grids for the `.vox` exporter tests (shared by the reference-writer pin, the golden hashes and the GPU export test)."""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def vox_cases():
    """Grids that exercise the MagicaVoxel writer: one cube, several 126^3 cubes in unusual first-seen orders, >255 labels
    (colour truncation), boundary-masked words, an empty export, and low-x-empty grids that hit the minCube* quirks."""
    rs = np.random.RandomState(11)
    cases = []
    g = rs.randint(0, 6, size=(20, 17, 30)).astype(np.uint16)
    cases.append(("small", g))
    g = np.zeros((130, 40, 140), np.uint16)
    g[rs.randint(0, 130, 4000), rs.randint(0, 40, 4000), rs.randint(0, 140, 4000)] = rs.randint(2, 700, 4000)
    g[5, 5, 130] = 0x8000 | 7  # boundary-masked word
    cases.append(("sparse_multi_cube", g))
    g = np.zeros((140, 130, 12), np.uint16)
    g[128:, :, :] = 3
    g[129, 127:, 4] = 300
    cases.append(("low_x_empty", g))
    g = np.zeros((140, 4, 140), np.uint16)
    g[130:, :, :] = 2
    cases.append(("low_x_empty_two_z_cubes", g))
    cases.append(("nothing_to_write", np.ones((9, 8, 7), np.uint16)))
    return cases


def labelled_fixture():
    """The reference's AL_12B occupancy fixture with a deterministic block labelling (ids 2..41)."""
    occ = np.load(os.path.join(HERE, "golden", "AL_12B_grid_128r.npz"))["grid"][:, :110, :].astype(np.uint16)
    x, y, z = np.meshgrid(np.arange(128), np.arange(110), np.arange(128), indexing="ij")
    return np.ascontiguousarray(occ * (2 + (x // 16 + 3 * (y // 16) + 5 * (z // 16)) % 40)).astype(np.uint16)


def all_vox_cases():
    return vox_cases() + [("AL_12B_blocks", labelled_fixture())]


def qstack_cases():
    """Grids for the `.qstack` exporter: noise (1x1 leaves), block labellings (uniform regions + bottom-up merges), columns with
    more than 255 runs (the uint8 interval counts wrap), the 0xFFFF wildcard value, odd dims, one-cell grids."""
    rs = np.random.RandomState(17)
    cases = [("one_cell", np.full((1, 1, 1), 5, np.uint16))]
    cases.append(("noise", (rs.randint(0, 5, size=(13, 9, 11)) * (rs.rand(13, 9, 11) < 0.6)).astype(np.uint16)))
    x, y, z = np.meshgrid(np.arange(40), np.arange(33), np.arange(20), indexing="ij")
    cases.append(("blocks", (2 + (x // 8 + 2 * (y // 8) + 3 * (z // 5)) % 7).astype(np.uint16)))
    g = np.zeros((16, 16, 24), np.uint16)
    g[:, :, 4:20] = 1
    g[3:9, 5:12, 8:15] = 4
    g[10:, :, 10:12] = 0xFFFF  # wildcard value: never merged upwards
    cases.append(("layers_with_wildcard", g))
    g = np.tile((np.arange(700) % 2 + 2).astype(np.uint16), (3, 2, 1))
    g[1, 1, ::5] = 9
    g[2, 0, :300] = 2
    cases.append(("more_than_255_runs", np.ascontiguousarray(g)))
    cases.append(("uniform", np.full((7, 5, 3), 2, np.uint16)))
    return cases


def all_qstack_cases():
    return qstack_cases() + [("AL_12B_blocks", labelled_fixture())]
