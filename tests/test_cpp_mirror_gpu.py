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


@pytest.mark.gpu
def test_cpp_mirror_generate_dataset(orc, tmp_path):
    """voxfrag::generateDataset (the call main.cpp:37-39 makes) over a folder with one .obj: files == the oracle replay."""
    from test_dataset_gpu import _replay

    import voxelfragmentml_b200 as vf
    from voxelfragmentml_b200 import dataset, synth

    _build()
    src = tmp_path / "in"
    src.mkdir()
    v, f = synth.vessel_mesh(3, n_ang=30, n_prof=14)
    with open(src / "VS_77.obj", "w") as fh:
        fh.write("".join(f"v {float(a)!r} {float(b)!r} {float(c)!r}\n" for a, b, c in v))
        fh.write("".join(f"f {a + 1} {b + 1} {c + 1}\n" for a, b, c in f))
    dest = str(tmp_path / "out") + "/"
    p = subprocess.run([BIN, "dataset", str(src), dest, "36"], capture_output=True, text=True, check=True)
    proc = vf.FragmentationProcedure(_fragmentInterval=(2, 3), _iterationInterval=(2, 1))
    proc._fractureParameters._clampVoxelMetricUnit = 36
    proc._fractureParameters._voxelPerMetricUnit = 36
    v2, f2 = dataset.load_obj(str(src / "VS_77.obj"))
    files, rows, generated, fragmentations, md, dims = _replay(orc, "VS_77", v2, f2, proc, orc.Rng(80))
    for rel, want in files.items():
        assert open(os.path.join(dest, rel), "rb").read() == want, rel
    assert p.stdout.split() == ["models", "1", "fragmentations", str(fragmentations), "fragments", str(generated), "files", str(len(files) + 3)]
