"""Builds libcdpr_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# CDPR_B200_LIB selects another build of the same library (kernel tuning experiments only)
LIB = os.environ.get("CDPR_B200_LIB") or os.path.join(HERE, "libcdpr_b200.so")
SOURCES = ["api.cu"]
HEADERS = ["common.cuh", "physics.cuh", "step_fast.cuh", "step_general.cuh", "misc_kernels.cuh", "../../include/cdpr_b200.h"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-shared"]


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.exists(os.path.join(CSRC, f)) and os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    if force or stale():
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        extra = os.environ.get("CDPR_NVCC_EXTRA", "").split()
        cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
        if verbose:
            print(r.stderr)
    return LIB


HOST_DIR = os.path.join(HERE, "host")
HOST_BIN = os.path.join(HOST_DIR, "cdpr_sinevelocitytest")


def build_host(force: bool = False) -> str:
    """The plugin-shaped C++ host shim + the headless sinevelocitytest driver, linked against the C ABI only."""
    srcs = [os.path.join(HOST_DIR, f) for f in ("CdprBatchPlugin.cpp", "sinevelocitytest_main.cpp")]
    deps = srcs + [os.path.join(HOST_DIR, "CdprBatchPlugin.h"), build()]
    if force or not os.path.exists(HOST_BIN) or any(os.path.getmtime(d) > os.path.getmtime(HOST_BIN) for d in deps):
        cmd = ["g++", "-O2", "-std=c++17", "-o", HOST_BIN] + srcs + ["-L" + HERE, "-lcdpr_b200", "-Wl,-rpath,$ORIGIN/.."]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("host shim build failed:\n" + r.stdout + r.stderr)
    return HOST_BIN


if __name__ == "__main__":
    import sys
    print(build(force=True, verbose="-v" in sys.argv))
    print(build_host(force=True))
