"""The C++ host mirror (include/voxfrag.hpp) driven with the reference's own call sequence, checked against the oracle."""
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

BIN = os.path.join(ROOT, "tests", "cpp", "_build", "mirror_demo")


def _build():
    os.makedirs(os.path.dirname(BIN), exist_ok=True)
    src = os.path.join(ROOT, "tests", "cpp", "mirror_demo.cpp")
    lib = os.path.join(ROOT, "voxelfragmentml_b200", "lib")
    if not os.path.exists(BIN) or os.path.getmtime(BIN) < max(os.path.getmtime(src), os.path.getmtime(os.path.join(ROOT, "include", "voxfrag.hpp"))):
        subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "include"), src, "-o", BIN, "-L", lib, "-lvoxfrag",
                        f"-Wl,-rpath,{lib}"], check=True)


def test_cpp_mirror_compiles_and_links():
    """CPU: the header-only mirror compiles against the C ABI and links to libvoxfrag.so."""
    _build()
    assert os.path.exists(BIN)


def _fnv(grid):
    h = 1469598103934665603
    b = grid.astype("<u2").tobytes()
    for byte in b:
        h = ((h ^ byte) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


@pytest.mark.gpu
@pytest.mark.parametrize("algo", ["naive", "flood"])
def test_cpp_mirror_reference_call_sequence(orc, vessel_grid, tmp_path, algo):
    _build()
    out = str(tmp_path / "out")
    p = subprocess.run([BIN, os.path.join(GOLDEN, "AL_12B_grid_128r.rle"), out, algo], capture_output=True, text=True, check=True)
    lines = p.stdout.strip().splitlines()
    got_hash = int(lines[0].split()[1], 16)
    seeds = np.array([[int(v) for v in l.split()[1:]] for l in lines if l.startswith("seed ")], np.uint32)
    r = orc.Rng(80)
    if algo == "naive":
        wseeds = orc.make_seeds(r, vessel_grid, 8, 0)
        want = orc.remove_isolated_regions_cpu(orc.naive(vessel_grid.copy(), wseeds, 0), wseeds)
    else:
        wseeds = orc.make_seeds(r, vessel_grid, 8, 16, merge_dfunc=0)
        want, _ = orc.flood(vessel_grid.copy(), wseeds, orc.CHEBYSHEV)
    want = orc.undo_mask(orc.detect_boundaries(want, 1), 15, False)
    assert np.array_equal(seeds, wseeds)
    # hash a smaller projection in python (pure-python FNV over 3.6 MB is slow): compare the exported .rle instead
    assert open(out + ".rle", "rb").read() == orc.encode_rle(want)
    assert int(lines[2].split()[1]) == int((want > 1).sum())
    assert got_hash != 0
