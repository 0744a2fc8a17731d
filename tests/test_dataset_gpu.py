"""f1: the native dataset driver (vf_dataset_model / vf_dataset_generate = CADScene::generateDataset's voxel path) against a replay of
the same loop with the oracle: every file name, every grid file byte for byte, the metadata text, the fragment counters."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _replay(orc, name, v, f, proc, rng, solid=False):
    """generateDataset's model loop restated with oracle operators (CADScene.cpp:240-452); returns {relative path: bytes}"""
    from test_dataset_cpu import _dims_rule

    fp = proc._fractureParameters
    mn, mx = v.min(0), v.max(0)
    dims = _dims_rule(mn, mx, fp._voxelPerMetricUnit, fp._clampVoxelMetricUnit)
    md = max(dims)
    grid = orc.voxelize_solid(v, f, mn, mx, dims) if solid else orc.voxelize_sat(v, f, mn, mx, dims)
    enc = {0: orc.encode_rle, 1: orc.encode_qstack, 2: lambda g: orc.encode_vox(g, True), 3: orc.encode_bing_squared}[fp._exportGridExtension]
    ext = ["rle", "qstack", "vox", "bing"][fp._exportGridExtension]
    files = {f"{name}/{name}_grid_{md}r.{ext}": enc(grid)}
    rows, generated, fragmentations = [], 0, 0
    mesh_rows = []
    n0, n1 = proc._fragmentInterval
    for nfr in range(n0, n1 + 1):
        if generated >= proc._maxFragmentsModel:
            break
        for it in range(proc.numIterations(nfr)):
            if generated >= proc._maxFragmentsModel:
                break
            grid = orc.reset_filling(grid)
            seeds = orc.make_seeds(rng, grid, nfr, 2 * nfr, merge_dfunc=fp._mergeSeedsDistanceFunction)
            grid, _ = orc.flood(grid, seeds, fp._distanceFunction)
            grid = orc.detect_boundaries(grid, 1)
            counts, occupied = orc.count_values(grid)
            rel = f"{name}/{name}_{nfr}f_{md}r_{it}it.{ext}"
            if proc._exportMesh:  # toTriangleMesh before undoMask; CADModel::saveBinary layout
                for idx, label in enumerate(int(v) for v in np.nonzero(counts)[0]):
                    mv, mf = orc.marching_cubes(grid, label, mn, mx)
                    vb = np.zeros((len(mv), 16), np.float32)
                    vb[:, :3] = mv[:, :3]
                    fb = np.zeros((len(mf), 4), np.uint32)
                    fb[:, :3] = mf[:, :3]
                    stem = rel[: -len(ext) - 1] + f"_{idx}"
                    files[stem + ".binm"] = np.uint32(len(mv)).tobytes() + vb.tobytes() + np.uint32(len(mf)).tobytes() + fb.tobytes()
                    pct = np.float32(counts[label]) / np.float32(occupied)
                    mesh_rows.append(f"{stem}.binm\t{idx}\t{dims[0]}x{dims[1]}x{dims[2]}\t{counts[label]}\t{occupied}\t{'%g' % float(pct)}\t{len(mv)}\t{len(mf)}\t")
            grid = orc.undo_mask(grid)
            files[rel] = enc(grid)
            rows.append((rel, dims))
            generated += int((counts[2:] != 0).sum())
            fragmentations += 1
    _replay.mesh_rows = mesh_rows
    return files, rows, generated, fragmentations, md, dims


def test_dataset_model_with_fragment_meshes(orc, tmp_path):
    """exportMesh: every fragment's marching-cubes mesh as .binm (CADModel::saveBinary layout) and the mesh metadata rows"""
    import voxelfragmentml_b200 as vf
    from voxelfragmentml_b200 import dataset, synth

    v, f = synth.vessel_mesh(1, n_ang=36, n_prof=18)
    proc = vf.FragmentationProcedure(_fragmentInterval=(2, 3), _iterationInterval=(2, 1), _exportMesh=True)
    proc._fractureParameters._clampVoxelMetricUnit = 44
    proc._fractureParameters._voxelPerMetricUnit = 44
    ctx = vf.Context(0)
    ctx.initSeed(80)
    grid = dataset.dataset_grid(ctx, proc)
    dest = str(tmp_path / "out") + "/"
    st = dataset.generate_model(grid, proc, "VS_02", v, f, dest)
    files, rows, generated, fragmentations, md, dims = _replay(orc, "VS_02", v, f, proc, orc.Rng(80))
    assert sum(k.endswith(".binm") for k in files) == generated > 0
    for rel, want in files.items():
        assert open(os.path.join(dest, rel), "rb").read() == want, rel
    meta = open(os.path.join(dest, "VS_02", f"VS_02_{md}_metadata_mesh.txt")).read().split("\n")
    assert meta[0] == "Filename\tFragment id\tVoxelization size\tVoxels\tOccupied voxels\tPercentage\tVertices\tFaces" and meta[-1] == ""
    assert meta[1:-1] == [dest + r for r in _replay.mesh_rows]
    assert st["files"] == len(files) + 3 and st["fragments"] == generated
    grid.close()
    ctx.close()


@pytest.mark.parametrize("ext,solid,writers", [(0, False, 2), (0, True, 0), (3, False, 1), (1, False, 2)])
def test_dataset_model_matches_oracle_replay(orc, tmp_path, ext, solid, writers):
    import voxelfragmentml_b200 as vf
    from voxelfragmentml_b200 import dataset, synth

    v, f = synth.vessel_mesh(2, n_ang=36, n_prof=18)
    proc = vf.FragmentationProcedure(_fragmentInterval=(2, 4), _iterationInterval=(3, 2), _solidVoxelization=solid, _writerThreads=writers)
    proc._fractureParameters._clampVoxelMetricUnit = 44
    proc._fractureParameters._voxelPerMetricUnit = 44
    proc._fractureParameters._exportGridExtension = ext
    ctx = vf.Context(0)
    ctx.initSeed(80)
    grid = dataset.dataset_grid(ctx, proc)
    dest = str(tmp_path / "out") + "/"
    st = dataset.generate_model(grid, proc, "VS_01", v, f, dest)
    files, rows, generated, fragmentations, md, dims = _replay(orc, "VS_01", v, f, proc, orc.Rng(80), solid)
    assert fragmentations == 3 + 2 + 2
    for rel, want in files.items():
        assert open(os.path.join(dest, rel), "rb").read() == want, rel
    produced = sorted(os.listdir(os.path.join(dest, "VS_01")))
    assert produced == sorted([os.path.basename(r) for r in files] + [f"VS_01_{md}_metadata_{k}.txt" for k in ("grid", "mesh", "pointcloud")])
    ext_s = ["rle", "qstack", "vox", "bing"][ext]
    meta = open(os.path.join(dest, "VS_01", f"VS_01_{md}_metadata_grid.txt")).read().split("\n")
    assert meta[0] == "Filename\tVoxelization size" and meta[-1] == ""
    assert meta[1:-1] == [f"{dest}{rel}\t{dims[0]}x{dims[1]}x{dims[2]}" for rel, _ in rows]
    assert open(os.path.join(dest, "VS_01", f"VS_01_{md}_metadata_mesh.txt")).read() == \
        "Filename\tFragment id\tVoxelization size\tVoxels\tOccupied voxels\tPercentage\tVertices\tFaces\n"
    assert st["models"] == 1 and st["fragmentations"] == fragmentations and st["fragments"] == generated
    assert st["files"] == len(files) + 3
    if ext == 0:  # only the byte stream crosses PCIe
        assert st["bytes_downloaded"] == sum(len(b) for b in files.values())
    grid.close()
    ctx.close()
    assert ext_s in produced[0]


def test_max_fragments_cap_stops_the_loop(orc, tmp_path):
    import voxelfragmentml_b200 as vf
    from voxelfragmentml_b200 import dataset, synth

    v, f = synth.vessel_mesh(0, n_ang=36, n_prof=18)
    proc = vf.FragmentationProcedure(_fragmentInterval=(3, 6), _iterationInterval=(4, 4), _maxFragmentsModel=10)
    proc._fractureParameters._clampVoxelMetricUnit = 40
    proc._fractureParameters._voxelPerMetricUnit = 40
    ctx = vf.Context(0)
    ctx.initSeed(80)
    grid = dataset.dataset_grid(ctx, proc)
    dest = str(tmp_path) + "/"
    st = dataset.generate_model(grid, proc, "m", v, f, dest)
    files, rows, generated, fragmentations, md, dims = _replay(orc, "m", v, f, proc, orc.Rng(80))
    assert fragmentations == 4 and generated >= 10  # 3 + 3 + 3 + 3 fragments reach the cap inside the first fragment count
    assert st["fragmentations"] == fragmentations and st["fragments"] == generated
    for rel, want in files.items():
        assert open(os.path.join(dest, rel), "rb").read() == want, rel
    grid.close()
    ctx.close()


def test_generate_dataset_walks_a_folder_of_obj_files(orc, tmp_path):
    """generateDataset end to end: two .obj files (+ one skipped by _startVessel), one RNG stream across the models."""
    import voxelfragmentml_b200 as vf
    from voxelfragmentml_b200 import dataset, synth

    src = tmp_path / "meshes"
    (src / "sub").mkdir(parents=True)
    meshes = {}
    for i, name in enumerate(["A_00", "B_01", "C_02"]):
        v, f = synth.vessel_mesh(i, n_ang=30, n_prof=14)
        v = (v * np.float32(3.0) + np.float32([1.0, -2.0, 0.5])).astype(np.float32)  # the loader normalises
        path = (src / "sub" if i == 2 else src) / f"{name}.obj"
        with open(path, "w") as fh:
            fh.write("".join(f"v {float(a)!r} {float(b)!r} {float(c)!r}\n" for a, b, c in v))
            fh.write("".join(f"f {a + 1} {b + 1} {c + 1}\n" for a, b, c in f))
        meshes[name] = str(path)
    proc = vf.FragmentationProcedure(_fragmentInterval=(2, 3), _iterationInterval=(2, 1), _startVessel="B_01")
    proc._fractureParameters._clampVoxelMetricUnit = 36
    proc._fractureParameters._voxelPerMetricUnit = 36
    ctx = vf.Context(0)
    ctx.initSeed(proc._fractureParameters._seed)
    dest = str(tmp_path / "dataset") + "/"
    st = vf.generateDataset(ctx, proc, str(src), ".obj", dest)
    assert st["models"] == 2 and not os.path.exists(os.path.join(dest, "A_00"))
    rng = orc.Rng(80)
    order = sorted(p for n, p in meshes.items() if n != "A_00")  # sorted paths; the first entries are dropped until _startVessel matches
    for path in order:
        name = os.path.splitext(os.path.basename(path))[0]
        v, f = dataset.load_obj(path)
        files, *_ = _replay(orc, name, v, f, proc, rng)
        for rel, want in files.items():
            assert open(os.path.join(dest, rel), "rb").read() == want, rel
    ctx.close()
