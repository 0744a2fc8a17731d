"""The "descent certificate" formulation of C1 that csrc/c1_descent.cu implements (the default path of vf_remove_isolated_regions), validated as an
ALGORITHM on the CPU: tools/c1_descent_prototype.py against the oracle's removeIsolatedRegionsCPU restatement (NaiveFracturer.cpp:111-150)."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, pick_seeds, random_blob_grid

sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.mark.parametrize("case", range(8))
def test_descent_certificate_equals_connected_to_seed_cleanup(orc, case):
    from c1_descent_prototype import c1_descent

    rs = np.random.RandomState(case)
    if case < 3:  # dense Voronoi labels under the three metrics: a handful of failing cells
        g = np.ones((40, 36, 48), np.uint16)
        seeds = pick_seeds(g, 24, case)
        lab = orc.naive(g.copy(), seeds, case)
    else:  # porous blobs: islands, long dead ends, stray FREE cells
        g = random_blob_grid((30, 28, 40), case, fill=float(rs.uniform(0.45, 0.7)), smooth=1)
        seeds = pick_seeds(g, 7, case)
        lab = orc.naive(g.copy(), seeds, case % 3)
        lab[1:3, 1:3, 1:3] = 1
    if case == 5:  # labels that lost their seed disappear
        seeds = seeds[:4]
    if case == 6:  # a seed planted on a cell that another seed labelled
        seeds = seeds.copy()
        seeds[0, :3] = np.argwhere(lab == seeds[1, 3])[0]
    if case == 7:  # a later seed takes the cell; the earlier one still starts its search there (CADScene's seed copies, CADScene.cpp:651)
        seeds = np.concatenate([seeds, [[seeds[2, 0], seeds[2, 1], seeds[2, 2], 77]]]).astype(np.uint32)
    want = orc.remove_isolated_regions_cpu(lab.copy(), seeds)
    got, nF, nD = c1_descent(lab, seeds)
    assert np.array_equal(got, want)
    assert nD >= nF
