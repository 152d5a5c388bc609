"""Build recipe for libcmwdense.so (the C-ABI library of hand-written sm_100a CUDA kernels).

``python -m cmw_rag_b200.build`` compiles every ``csrc/*.cu`` with
``nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo`` and links them in-tree into
``cmw_rag_b200/csrc/libcmwdense.so`` (git-ignored; it travels to the GPU box with the snapshot).
nvcc cross-compiles without a GPU, so this also is the "does it build" check.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(CSRC, "libcmwdense.so")
SOURCES = ["store.cu", "scan.cu", "gemm.cu", "gemm2.cu", "pool.cu", "multivector.cu", "exchange.cu", "api.cu"]
HEADERS = ["common.cuh", "ptx.cuh", "gemm_common.cuh", os.path.join("..", "..", "include", "cmw_dense.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: libcmwdense.so cannot be built (there is no CPU fallback)")
    return exe


def _mtime(path: str) -> float:
    return os.path.getmtime(path) if os.path.exists(path) else 0.0


def _compile(src: str, verbose: bool, extra: list[str]) -> str:
    obj = os.path.join(OBJ, os.path.splitext(src)[0] + ".o")
    cmd = [nvcc(), *NVCC_FLAGS, *extra, "-c", os.path.join(CSRC, src), "-o", obj]
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{res.stdout}\n{res.stderr}")
    if verbose and (res.stdout or res.stderr):
        print(res.stdout, res.stderr, file=sys.stderr)
    return obj


def build(force: bool = False, verbose: bool = False, ptxas_info: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    hdr_time = max(_mtime(os.path.join(CSRC, h)) for h in HEADERS)
    extra = ["-Xptxas", "-v"] if ptxas_info else []
    todo = []
    objs = []
    for src in SOURCES:
        obj = os.path.join(OBJ, os.path.splitext(src)[0] + ".o")
        objs.append(obj)
        if force or ptxas_info or _mtime(obj) < max(_mtime(os.path.join(CSRC, src)), hdr_time):
            todo.append(src)
    if todo:
        with cf.ThreadPoolExecutor(max_workers=min(len(todo), os.cpu_count() or 4)) as ex:
            list(ex.map(lambda s: _compile(s, verbose or ptxas_info, extra), todo))
    if todo or not os.path.exists(LIB):
        cmd = [nvcc(), "-shared", "-o", LIB, *objs, "-cudart", "static", "-Xcompiler", "-fPIC"]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True, ptxas_info="--ptxas" in sys.argv))
