"""bench.py's JSON contract: the reference arm runs on the host alone (CPU test); the CUDA arm is checked on the GPU box at a small size."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config"}


def _run(*args):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, check=True, cwd=ROOT)
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, p.stdout
    return json.loads(lines[0])


@pytest.mark.parametrize("ref_impl", ["port", "auto"])
def test_reference_arm_line(ref_impl):
    d = _run("--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-size", "64", "--ref-impl", ref_impl)
    assert BASE_KEYS <= set(d) and d["impl"] == "reference" and d["metric"] == "Gvoxels/s fragmented at 512^3" and d["unit"] == "Gvoxels/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    if ref_impl == "port":
        assert cb["kind"] == "port"
    elif all(os.path.exists(os.path.join(ROOT, "oracle", "_ref", f)) for f in ("libvf_ref.so", "libvf_ref_glsl.so")):
        assert cb["kind"] == "reference" and "shaders" in cb["sample"]
    assert d["sample_grid"] == [64, 64, 64] and d["same_size_as_config"] is False and d["config"]["grid"] == [512, 512, 512]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--cpu-size", "32"], capture_output=True, text=True,
                       env=env, cwd=ROOT)
    assert p.returncode == 0 and p.stdout.strip() == ""


@pytest.mark.gpu
def test_cuda_arm_line_small():
    d = _run("--steps", "2", "--warmup", "3", "--size", "128", "--cpu-size", "64")
    assert BASE_KEYS <= set(d) and "impl" not in d and d["metric"] == "Gvoxels/s fragmented at 512^3"
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3 and d["value"] > 0 and d["data"] == "synthetic"
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["peak"] > 0 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and "traffic" in r
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] == 128 ** 3 // 8 and 12 < e["d2h_bytes_per_step"] < 2 * 128 ** 3 and e["rle_stream_equals_grid"] is True
    f = e["full_grid"]
    assert f["value"] > 0 and f["h2d_bytes_per_step"] == 2 * 128 ** 3 and f["d2h_bytes_per_step"] >= 2 * 128 ** 3
    assert d["gpu_launches"] > 0 and {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["batch"]["unit"] == "models/s" and d["batch"]["value"] > 0
