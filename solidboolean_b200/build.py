"""Build the sm_100a shared library in-tree (nvcc cross-compiles without a GPU).

    python -m solidboolean_b200.build [--force] [--verbose]

Output: solidboolean_b200/lib/libsolidboolean_b200.so (git-ignored, travels to the
GPU box with the snapshot).  Flags of note:
  -gencode arch=compute_100a,code=sm_100a   B200 only, no other targets
  -fmad=false                               no FMA contraction anywhere (the
                                            predicate / classifier must round every
                                            binary64 op separately, like the reference)
  -lineinfo                                 so ncu's source page maps to the .cu files
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build")
LIB = os.path.join(LIBDIR, "libsolidboolean_b200.so")
SOURCES = ["sb_capi.cu", "sb_build.cu", "sb_broad.cu", "sb_narrow.cu", "sb_classify.cu", "sb_classify2.cu", "sb_grid.cu", "sb_halfedge.cu", "sb_cuts.cu", "sb_shard.cu", "sb_comm.cu", "sb_flood.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-fmad=false",
         "-Xcompiler", "-fPIC,-O2,-Wall,-Wno-unused-function", "-Xptxas", "-v"]


def _deps():
    out = []
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in os.listdir(root):
            if f.endswith((".cu", ".cuh", ".h")):
                out.append(os.path.join(root, f))
    return out


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    deps = _deps() + [os.path.abspath(__file__)]
    if not force and not _stale(LIB, deps):
        return LIB
    logs = {}

    def compile_one(src):
        obj = os.path.join(OBJDIR, src.replace(".cu", ".o"))
        cmd = [NVCC] + FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        logs[src] = r.stdout + r.stderr
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (src, logs[src]))
        return obj

    with cf.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    with open(os.path.join(OBJDIR, "ptxas.log"), "w") as f:
        for k, v in logs.items():
            f.write("==== %s ====\n%s\n" % (k, v))
    if verbose:
        for k, v in logs.items():
            print("====", k)
            print(v)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
