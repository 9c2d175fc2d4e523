"""Builds libiftwt.so (hand-written sm_100a CUDA + the C ABI) in-tree with nvcc.

No torch.utils.cpp_extension: the library has no torch types in its boundary and
is loaded with ctypes (or linked by a C++ host)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libiftwt.so")
SOURCES = ["grid.cu", "knn.cu", "graph.cu", "normals.cu", "seeds.cu", "ift.cu", "voxel.cu", "hierarchy.cu", "shard.cu",
           "capi.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    # parity: no FMA contraction anywhere (FP32 op order is part of the contract)
    "-fmad=false", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "-ccbin", "/usr/bin/g++",
]


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    headers = [os.path.join(CSRC, "common.cuh"), os.path.join(HERE, "..", "include", "iftwt.h")]
    objs = []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(LIBDIR, s.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, [src] + headers):
            cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            if verbose:
                print(" ".join(cmd), file=sys.stderr)
            subprocess.check_call(cmd)
    if force or _stale(LIB, objs):
        cmd = [NVCC, "-shared", "-ccbin", "/usr/bin/g++", "-o", LIB] + objs + ["-lcudart_static", "-ldl", "-lrt",
                                                                             "-lpthread"]
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
