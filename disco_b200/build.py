"""In-tree build of the native pieces (sm_100a only).  `python -m disco_b200.build` or __graft_entry__.build()."""
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
HOST = os.path.join(PKG, "host")
GPU_LIB = os.path.join(PKG, "libdisco_gpu.so")
HOST_LIB = os.path.join(PKG, "libdisco_host.so")
BUILDG = os.path.join(PKG, "bin", "buildG")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd, verbose):
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.run(cmd, check=True)


def build_gpu(force=False, verbose=False):
    srcs = [os.path.join(CSRC, f) for f in ("kernels.cu", "api.cu", "simplify.cu")]
    deps = srcs + [os.path.join(CSRC, f) for f in ("dna.cuh", "kernels.cuh", "edges_flat.cuh")] + [os.path.join(ROOT, "include", "disco_gpu.h")]
    if force or _newer(GPU_LIB, deps):
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        _run([nvcc] + NVCC_FLAGS + ["-o", GPU_LIB] + srcs, verbose)
    return GPU_LIB


def build_host(force=False, verbose=False):
    if not os.path.isdir(HOST):
        return None
    srcs = [os.path.join(HOST, f) for f in sorted(os.listdir(HOST)) if f.endswith(".cpp") and f != "buildg_main.cpp"]
    if not srcs:
        return None
    deps = srcs + [os.path.join(HOST, f) for f in os.listdir(HOST) if f.endswith(".h")] + \
        [os.path.join(ROOT, "include", f) for f in os.listdir(os.path.join(ROOT, "include"))]
    if force or _newer(HOST_LIB, deps):
        _run(["g++", "-O3", "-mpopcnt", "-std=c++17", "-fopenmp", "-fPIC", "-shared", "-Wall", "-o", HOST_LIB] + srcs + ["-lz"], verbose)
    main = os.path.join(HOST, "buildg_main.cpp")
    if os.path.exists(main) and (force or _newer(BUILDG, deps + [main, GPU_LIB])):
        os.makedirs(os.path.dirname(BUILDG), exist_ok=True)
        _run(["g++", "-O3", "-std=c++17", "-fopenmp", "-Wall", "-o", BUILDG, main, "-L" + PKG, "-ldisco_host", "-ldisco_gpu",
              "-Wl,-rpath,$ORIGIN/..", "-lz"], verbose)
    return HOST_LIB


def build_all(force=False, verbose=False):
    build_gpu(force, verbose)
    build_host(force, verbose)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose=True)
