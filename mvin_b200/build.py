"""Build libmvin_b200.so in-tree (mvin_b200/lib/) with nvcc for sm_100a.  `python -m mvin_b200.build`.

Translation units: mvin_capi.cu (the C ABI) and mvin_steps.cu compiled once per embedding dimension
(-DMVIN_DIM=8|16|32|64|128; the forward / backward orchestration and every kernel instantiated for that dimension),
all in parallel, then linked into one shared library."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
OBJ_DIR = os.path.join(LIB_DIR, "obj")
LIB_PATH = os.path.join(LIB_DIR, "libmvin_b200.so")
DIMS = [8, 16, 32, 64, 128]
SOURCES = ["mvin_capi.cu", "mvin_steps.cu"]
HEADERS = ["common.cuh", "gemm.cuh", "level.cuh", "level_tc.cuh", "level_tcb.cuh", "misc.cuh", "user.cuh", "umma.cuh", "umma_bf.cuh", "gemm_tc.cuh", "host.cuh",
           "steps.cuh", "table.cuh", "group.cuh", "exchange.cuh", os.path.join("..", "..", "include", "mvin_b200.h")]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC"]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libmvin_b200.so")


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _units():
    yield "mvin_capi.cu", os.path.join(OBJ_DIR, "capi.o"), []
    for d in DIMS:
        yield "mvin_steps.cu", os.path.join(OBJ_DIR, f"steps_{d}.o"), [f"-DMVIN_DIM={d}"]


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB_PATH
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = _nvcc()
    extra = ["-Xptxas", "-v"] if verbose else []

    def compile_one(unit):
        src, obj, defs = unit
        cmd = [nvcc] + NVCC_FLAGS + extra + defs + ["-c", src, "-o", obj]
        res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
        return cmd, res

    units = list(_units())
    with ThreadPoolExecutor(max_workers=min(len(units), os.cpu_count() or 1)) as pool:
        results = list(pool.map(compile_one, units))
    for cmd, res in results:
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
        if verbose:
            print(" ".join(cmd[-4:]))
            print(res.stderr)
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB_PATH] + [u[1] for u in units]
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
