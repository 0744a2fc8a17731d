"""In-tree build of libvoxfrag.so (hand-written CUDA for sm_100a behind the C ABI of include/voxfrag.h).

    python build_lib.py [--force] [--verbose]

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
HERE = os.path.join(ROOT, "voxelfragmentml_b200")
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libvoxfrag.so")
INCLUDE = os.path.join(ROOT, "include")

NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
HOSTCXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
# -fmad=false is applied per file where float32 operation order is part of the contract (voxelize.cu: SAT predicate)
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-ccbin", HOSTCXX, "-Xcompiler", "-fPIC,-O2,-Wall,-ffp-contract=off",
          "-I", INCLUDE, "--expt-relaxed-constexpr", "-Xptxas", "-v"]
PER_FILE = {"voxelize.cu": ["-fmad=false"]}
EXTRA = os.environ.get("VF_NVCC_EXTRA", "").split()  # e.g. -DVF_FLOOD_TIMING for the instrumented flood kernel (tools/ only)


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith((".cu", ".cpp")))


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src, verbose):
    out = os.path.join(OBJ, os.path.splitext(src)[0] + ".o")
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))] + [os.path.join(INCLUDE, "voxfrag.h"), __file__]
    path = os.path.join(CSRC, src)
    if not _stale(out, [path] + hdrs):
        return out, ""
    cmd = [NVCC, *ARCH, *COMMON, *PER_FILE.get(src, []), *EXTRA, "-x", "cu", "-c", path, "-o", out]
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{' '.join(cmd)}\n{p.stdout}\n{p.stderr}")
    return out, p.stderr


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    srcs = sources()
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(lambda s: _compile(s, verbose), srcs))
    objs = [r[0] for r in results]
    log = "".join(r[1] for r in results)
    if log:
        with open(os.path.join(OBJ, "ptxas.log"), "a") as f:
            f.write(log)
        if verbose:
            print(log)
    if _stale(LIB, objs):
        cmd = [NVCC, *ARCH, "-shared", "-ccbin", HOSTCXX, "-o", LIB, *objs, "-cudart", "static"]
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError(f"link failed:\n{' '.join(cmd)}\n{p.stdout}\n{p.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
