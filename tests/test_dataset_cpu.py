"""Host logic of the dataset driver (csrc/dataset.cpp) that needs no GPU: the metric dims rule, the iteration schedule, the .obj
reader.  Expected values are restated here from CADScene.cpp:262-273, :304-306 and CADModel.cpp:148-152 in numpy float32."""
import numpy as np


def _dims_rule(mn, mx, per_unit, clamp):
    size = (np.float32(mx) - np.float32(mn)).astype(np.float32)
    v = np.ceil(size * np.float32(per_unit)).astype(np.int64)
    if (v > clamp).any():
        v = np.floor((np.float32(clamp) * size) / size.max()).astype(np.int64)
    while v[0] % 4:
        v[0] += 1
    while v[2] % 4:
        v[2] += 1
    return tuple(int(a) for a in v)


def test_dataset_dims_rule_matches_the_reference_rule():
    from voxelfragmentml_b200.dataset import dataset_dims

    rs = np.random.RandomState(2)
    for _ in range(300):
        c = rs.uniform(-1, 1, 3)
        h = rs.uniform(0.05, 0.5, 3)
        mn, mx = np.float32(c - h), np.float32(c + h)
        per_unit, clamp = int(rs.choice([20, 64, 128, 200])), int(rs.choice([64, 128, 200, 256]))
        assert dataset_dims(mn, mx, per_unit, clamp) == _dims_rule(mn, mx, per_unit, clamp)
    # the normalised vessel at the reference's dataset defaults (200 per unit, clamp 200)
    mn, mx = np.float32([-0.3428, -0.499999, -0.3428]), np.float32([0.3428, 0.499999, 0.3428])
    assert dataset_dims(mn, mx, 200, 200) == (140, 200, 140)


def test_iteration_schedule_is_glm_mix_of_the_interval():
    import voxelfragmentml_b200 as vf

    p = vf.FragmentationProcedure()
    assert (p._fragmentInterval, p._iterationInterval, p._maxFragmentsModel) == ((2, 10), (25, 15), 1000)
    fp = p._fractureParameters
    assert (fp._erode, fp._biasSeeds, fp._voxelPerMetricUnit, fp._exportGridExtension) == (0, 0, fp._clampVoxelMetricUnit, vf.ExportGrid.RLE)
    want = []
    for n in range(2, 11):
        a = np.float32(n - 2) / np.float32(8)
        want.append(int(np.float32(25) * (np.float32(1) - a) + np.float32(15) * a))
    assert [p.numIterations(n) for n in range(2, 11)] == want == [25, 23, 22, 21, 20, 18, 17, 16, 15]
    q = vf.FragmentationProcedure(_fragmentInterval=(4, 4), _iterationInterval=(3, 9))
    assert q.numIterations(4) == 3


def test_obj_reader_triangulates_and_normalises(tmp_path):
    from voxelfragmentml_b200.dataset import load_obj

    path = tmp_path / "box.obj"
    path.write_text("# a quad, a triangle with texture/normal indices, a relative-index face\n"
                    "v 1 2 3\nv 5 2 3\nv 5 4 3\nv 1 4 3\nv 3 3 11\nvn 0 0 1\nvt 0 0\n"
                    "f 1 2 3 4\nf 1/1/1 2/1/1 5/1/1\nf -1 -2 -3\n")
    v, f = load_obj(str(path))
    assert f.tolist() == [[0, 1, 2], [0, 2, 3], [0, 1, 4], [4, 3, 2]]
    raw = np.float32([[1, 2, 3], [5, 2, 3], [5, 4, 3], [1, 4, 3], [3, 3, 11]])
    mn, mx = raw.min(0), raw.max(0)
    ctr = (mx + mn) / np.float32(2)
    s = np.float32(0.499999) / (mx - ctr).max() * np.float32(2)
    want = (s * raw + s * (-ctr)).astype(np.float32)
    assert np.array_equal(v, want)
    assert abs(float(v[:, 2].max()) - 0.999998) < 1e-6  # the longest axis spans 2 * 0.999998
