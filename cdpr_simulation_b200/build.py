"""Builds libcdpr_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# CDPR_B200_LIB selects another build of the same library (kernel tuning experiments only)
LIB = os.environ.get("CDPR_B200_LIB") or os.path.join(HERE, "libcdpr_b200.so")
# one translation unit per group of kernel instances: they compile in parallel, then link into the one .so
SOURCES = ["api.cu", "comm.cu", "general.cu", "flex.cu", "flex_u2.cu", "flex_u4.cu", "flexr.cu", "flexr_nc8l2.cu", "flexr_nc4l1.cu", "flexr_nc8l1.cu", "fast_nc4_base.cu", "fast_nc4_diag.cu", "fast_nc4_spec.cu", "fast_nc8_base.cu", "fast_nc8_diag.cu",
           "fast_nc8_spec.cu", "fast_nc8_pair.cu"]
HEADERS = ["common.cuh", "physics.cuh", "step_fast.cuh", "step_general.cuh", "step_flex.cuh", "step_flexr.cuh", "flexr_common.cuh", "legs.cuh", "misc_kernels.cuh", "launch.h", "fast_inst.cuh",
           "../../include/cdpr_b200.h"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"]
OBJ_DIR = os.path.join(HERE, "build")


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.exists(os.path.join(CSRC, f)) and os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def _compile_one(args):
    cmd, src = args
    r = subprocess.run(cmd, capture_output=True, text=True)
    return src, r.returncode, r.stdout + r.stderr


def build(force: bool = False, verbose: bool = False) -> str:
    """nvcc -c every unit that is older than its sources (in parallel), then one nvcc -shared link."""
    if not (force or stale()):
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("CDPR_NVCC_EXTRA", "").split()
    tag = ("_" + "".join(c if c.isalnum() else "_" for c in " ".join(extra))) if extra else ""
    os.makedirs(OBJ_DIR, exist_ok=True)
    newest_header = max(os.path.getmtime(os.path.join(CSRC, f)) for f in HEADERS)
    jobs, objs = [], []
    for s in SOURCES:
        src, obj = os.path.join(CSRC, s), os.path.join(OBJ_DIR, s[:-3] + tag + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(newest_header, os.path.getmtime(src)):
            jobs.append(([nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src], s))
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        for src, rc, out in ex.map(_compile_one, jobs):
            if rc != 0:
                raise RuntimeError(f"nvcc failed on {src}:\n{out}")
            if verbose:
                print(out)
    r = subprocess.run([nvcc, "-shared", "-o", LIB] + objs, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return LIB


HOST_DIR = os.path.join(HERE, "host")
HOST_BIN = os.path.join(HOST_DIR, "cdpr_sinevelocitytest")
HOST_MULTI_BIN = os.path.join(HOST_DIR, "cdpr_multigpu_check")   # configs 4 and 5 from a C++ host, one process, G devices


def build_host(force: bool = False) -> str:
    """The plugin-shaped C++ host shim + the headless sinevelocitytest driver, linked against the C ABI only."""
    srcs = [os.path.join(HOST_DIR, f) for f in ("CdprBatchPlugin.cpp", "sinevelocitytest_main.cpp")]
    deps = srcs + [os.path.join(HOST_DIR, "CdprBatchPlugin.h"), build()]
    if force or not os.path.exists(HOST_BIN) or any(os.path.getmtime(d) > os.path.getmtime(HOST_BIN) for d in deps):
        cmd = ["g++", "-O2", "-std=c++17", "-o", HOST_BIN] + srcs + ["-L" + HERE, "-lcdpr_b200", "-Wl,-rpath,$ORIGIN/.."]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("host shim build failed:\n" + r.stdout + r.stderr)
    multi = os.path.join(HOST_DIR, "multi_gpu_main.cpp")
    if force or not os.path.exists(HOST_MULTI_BIN) or any(os.path.getmtime(d) > os.path.getmtime(HOST_MULTI_BIN) for d in (multi, build())):
        cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
        cmd = ["g++", "-O2", "-std=c++17", "-o", HOST_MULTI_BIN, multi, "-I" + os.path.join(cuda, "include"), "-L" + HERE, "-lcdpr_b200",
               "-L" + os.path.join(cuda, "lib64"), "-lcudart", "-Wl,-rpath,$ORIGIN/..", "-Wl,-rpath," + os.path.join(cuda, "lib64")]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("multi-GPU host driver build failed:\n" + r.stdout + r.stderr)
    return HOST_BIN


if __name__ == "__main__":
    import sys
    print(build(force=True, verbose="-v" in sys.argv))
    print(build_host(force=True))
