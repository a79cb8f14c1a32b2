"""Build libnlk.so (hand-written sm_100a CUDA + host C++) in-tree with nvcc.  No JIT, no torch extension:
the .so is a plain C-ABI library (include/nlk.h) and travels to the GPU box with the repo snapshot."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libnlk.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
SOURCES = ["nlk_basis.cpp", "nlk_mesh.cpp", "nlk_dense.cpp", "nlk_kernels.cu", "nlk_kernels_tp.cu", "nlk_solver.cu", "nlk_api.cu", "nlk_krylov.cu", "nlk_coarse.cu", "nlk_cgp.cu", "nlk_schwarz.cu"]
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC,-fopenmp,-O3", "-Xcudafe", "--diag_suppress=177"]


def _stale(src, obj):
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    deps = [src] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".hpp", ".cuh", ".h"))]
    deps.append(os.path.join(HERE, "..", "include", "nlk.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    jobs = []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ, s.rsplit(".", 1)[0] + ".o")
        if force or _stale(src, obj):
            cmd = [NVCC] + FLAGS + (["-x", "cu"] if s.endswith(".cpp") else []) + ["-c", src, "-o", obj]
            jobs.append(cmd)
    def run(cmd):
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        return r
    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(run, jobs))
    objs = [os.path.join(OBJ, s.rsplit(".", 1)[0] + ".o") for s in SOURCES]
    if jobs or not os.path.exists(LIB):
        run([NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fopenmp",
                                                    "-lcudart", "-ldl", "-lgomp"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
