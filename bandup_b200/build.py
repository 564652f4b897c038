"""Builds libbandup_gpu.so (sm_100a only) in-tree with nvcc.

    python -m bandup_b200.build [--force]

The shared library lands in bandup_b200/lib/ so that it travels with the repo
snapshot to the GPU box; it is git-ignored (*.so).
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIBDIR = PKG / "lib"
LIB = LIBDIR / "libbandup_gpu.so"
HOST_SRC = PKG / "host" / "bandup_unfold_gpu.cpp"
HOST_EXE = LIBDIR / "BandUP_gpu.x"   # the native (C++) host above the C ABI: `bandup unfold` without symmetry
SOURCES = ["capi.cu", "capi_projected.cu", "capi_symm.cu", "coset_classify.cu", "weights_reduce.cu", "delta_n.cu", "rho.cu", "symm_average.cu", "gvecs.cu", "fold_map.cu", "comm.cu", "synth.cu"]
HEADERS = [CSRC / "kernels.cuh", CSRC / "ctx.cuh", CSRC / "nccl_dyn.h", CSRC / "stage_pipeline.h", PKG.parent / "include" / "bandup_gpu.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O2,-Wall",
    "-Xptxas", "-v",
    "--shared", "-cudart", "static",
]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(nvcc).exists():
        raise RuntimeError("nvcc not found: libbandup_gpu.so cannot be built (no CPU fallback exists)")
    return nvcc


def _digest() -> str:
    h = hashlib.sha256()
    for p in [CSRC / s for s in SOURCES] + HEADERS + [HOST_SRC]:
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    LIBDIR.mkdir(exist_ok=True)
    stamp = LIBDIR / "libbandup_gpu.digest"
    digest = _digest()
    if LIB.exists() and HOST_EXE.exists() and stamp.exists() and stamp.read_text() == digest and not force:
        return LIB
    env = dict(os.environ)
    env.pop("CC", None)   # this image exports CC/CXX pointing at a gcc without libgomp specs
    env.pop("CXX", None)
    cmd = [_nvcc(), *NVCC_FLAGS, "-ccbin", shutil.which("g++") or "g++",
           "-I", str(PKG.parent / "include"),
           "-o", str(LIB), *[str(CSRC / s) for s in SOURCES], "-ldl"]
    res = subprocess.run(cmd, capture_output=True, text=True, env=env)
    (LIBDIR / "build.log").write_text(" ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libbandup_gpu.so (see bandup_b200/lib/build.log)")
    host = [shutil.which("g++") or "g++", "-O2", "-std=c++17", "-Wall", "-pthread", "-I", str(PKG.parent / "include"),
            str(HOST_SRC), "-L", str(LIBDIR), "-lbandup_gpu", "-Wl,-rpath,$ORIGIN", "-o", str(HOST_EXE)]
    res = subprocess.run(host, capture_output=True, text=True, env=env)
    with open(LIBDIR / "build.log", "a") as f:
        f.write(" ".join(host) + "\n" + res.stdout + res.stderr)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("g++ failed building BandUP_gpu.x (see bandup_b200/lib/build.log)")
    stamp.write_text(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
