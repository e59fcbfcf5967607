"""In-tree build of the CUDA extension: nvcc -> values_b200/lib/libvalues_b200.so (sm_100a).

The .so is a plain C-ABI shared library (include/values_b200.h); Python binds it with
ctypes (values_b200/_lib.py).  Built artefacts are git-ignored but travel to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB_DIR = os.path.join(PKG, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libvalues_b200.so")
OBJ_DIR = os.path.join(PKG, "build")
SOURCES = ["api.cu", "uncertainty.cu", "aggregate.cu", "stitch.cu", "stats.cu", "formats.cu"]
HEADERS = [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "tma_host.cuh"), os.path.join(CSRC, "log64_table.inc"),
           os.path.join(ROOT, "include", "values_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-I" + os.path.join(ROOT, "include"), "-I" + CSRC,
]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: the values_b200 CUDA extension cannot be built")
    return nvcc


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, sanitize: bool = False) -> str:
    """Compile every .cu for sm_100a and link the shared library.  Returns its path.
    sanitize=True builds values_b200/lib_sanitize/libvalues_b200.so with -DVALUES_ALL_LANES_ARRIVE (every
    lane of a consumer warp arrives on the ring's empty barrier itself, see csrc/common.cuh) for
    compute-sanitizer racecheck runs; VALUES_B200_LIB=<path> makes values_b200._lib load it."""
    global LIB_DIR, LIB_PATH, OBJ_DIR, NVCC_FLAGS
    if sanitize:
        LIB_DIR, OBJ_DIR = os.path.join(PKG, "lib_sanitize"), os.path.join(PKG, "build_sanitize")
        LIB_PATH = os.path.join(LIB_DIR, "libvalues_b200.so")
        NVCC_FLAGS = NVCC_FLAGS + ["-DVALUES_ALL_LANES_ARRIVE"]
    os.makedirs(LIB_DIR, exist_ok=True)
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = _nvcc()
    jobs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        if force or _stale(o, [s] + HEADERS):
            jobs.append([nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        return r.stderr

    with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as ex:
        for log in ex.map(run, jobs):
            if verbose and log:
                print(log, file=sys.stderr)
    objs = [os.path.join(OBJ_DIR, s.replace(".cu", ".o")) for s in SOURCES]
    if force or jobs or _stale(LIB_PATH, objs):
        run([nvcc, "-shared", "-o", LIB_PATH] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"])
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, sanitize="--sanitize" in sys.argv))
