"""Builds ``arvae_b200/csrc/libarvae_b200.so`` in-tree with nvcc for sm_100a.

The library has no torch dependency (C ABI only, see include/arvae_b200.h), so a
plain ``nvcc -shared`` is all that is needed; it cross-compiles on a box
without a GPU and the resulting ``.so`` travels with the repo snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.environ.get("ARVAE_LIB_OUT") or os.path.join(CSRC, "libarvae_b200.so")  # override: A/B experiments
SOURCES = ["api.cu", "reg_dense.cu", "reg_sorted.cu", "sort.cu", "latent_head.cu", "head_fused.cu", "music_attrs.cu", "eval_metrics.cu"]
HEADERS = ["common.cuh", "reg_internal.cuh", os.path.join("..", "..", "include", "arvae_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    # no --use_fast_math: the attribute compares must not flush subnormals
    "--ftz=false", "--prec-div=true", "--prec-sqrt=true", "--fmad=true",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libarvae_b200.so cannot be built")


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    deps += [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu to an object (in parallel) and link the shared library."""
    if not force and not _stale():
        return LIB
    nvcc = _nvcc()
    objdir = os.environ.get("ARVAE_OBJ_DIR") or os.path.join(CSRC, "build")
    os.makedirs(objdir, exist_ok=True)
    env = dict(os.environ)
    env.pop("CC", None)  # the image's CC points at a wrapper that breaks nvcc's host compile
    env.pop("CXX", None)

    def compile_one(src: str) -> str:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        extra = os.environ.get("ARVAE_NVCC_EXTRA", "").split()  # experiments only (e.g. -DARVAE_TILE_THREADS=256)
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", os.path.join(CSRC, src), "-o", obj]
        p = subprocess.run(cmd, capture_output=True, text=True, env=env)
        log = p.stdout + p.stderr
        with open(obj + ".log", "w") as f:
            f.write(" ".join(cmd) + "\n" + log)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{log}")
        if verbose:
            print(log, file=sys.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
           "-Xcompiler", "-fPIC", "-ldl"]  # -ldl: NVTX v3 loads its tool backend with dlopen
    p = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if p.returncode != 0:
        raise RuntimeError("link failed:\n" + p.stdout + p.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
