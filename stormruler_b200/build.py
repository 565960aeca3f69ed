"""Build libstormb200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m stormruler_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU. Objects are cached per source under csrc/_build/ (git-ignored);
the .so is git-ignored too but travels to the GPU box with the gpurun snapshot. Flags: -lineinfo so
ncu's source page maps to these files; --fmad=false and -ffp-contract=off because the floating-point
contract of the path forbids FMA contraction (every operation rounded separately, like the
reference's canonical g++ -O2 build).
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_build")
LIB = os.path.join(HERE, "libstormb200.so")
SOURCES = ["sb_api.cu", "sb_op.cu", "sb_solvers.cu", "sb_mega.cu", "sb_gmres.cu", "sb_comm.cu", "sb_mesh_host.cpp", "sb_part_host.cpp"]
METIS = "/usr/local/cuda/targets/x86_64-linux/lib/libmetis_static.a"  # ships with the CUDA toolkit
HEADERS = [os.path.join(CSRC, h) for h in ("sb_common.cuh", "sb_kernels.cuh", "sb_group_body.cuh", "sb_apply_rows.cuh", "sb_op.cuh", "sb_comm.cuh", "sb_solver_bodies.cuh", "sb_mega.cuh")] + \
    [os.path.join(HERE, "..", "include", "stormb200.h"), os.path.abspath(__file__)]

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-std=c++20", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "--fmad=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=hidden,-Wall",
]


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _compile(src: str, verbose: bool) -> str:
    obj = os.path.join(OBJ, os.path.splitext(src)[0] + ".o")
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}")
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    todo = [s for s in SOURCES
            if force or _stale(os.path.join(OBJ, os.path.splitext(s)[0] + ".o"), [os.path.join(CSRC, s)] + HEADERS)]
    with ThreadPoolExecutor(max_workers=8) as ex:
        list(ex.map(lambda s: _compile(s, verbose), todo))
    objs = [os.path.join(OBJ, os.path.splitext(s)[0] + ".o") for s in SOURCES]
    if todo or _stale(LIB, objs):
        res = subprocess.run([NVCC, "-shared", "-o", LIB] + objs + [METIS, "-lpthread", "-ldl"], capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("link of libstormb200.so failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
