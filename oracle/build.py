"""Build recipe for the oracle's C restatement and the HNSW CPU baseline.

TEST INFRASTRUCTURE ONLY.  Outputs go next to the sources as ``*.so`` (git-ignored,
not gpurun-ignored, so a prebuilt copy travels to the GPU box; both are rebuilt on
demand when missing because the box has the same gcc).

The reference itself is pure Python with its vector arithmetic inside the external
``chromadb==1.3.0`` server (rag_engine/requirements_frozen.txt:15), so there is no
reference source to compile into ``oracle/_ref`` -- see DESIGN.md "Oracle".
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))

# x86-64-v3 baseline + runtime-dispatched AVX-512 clones (target_clones in the sources):
# the GPU box's host CPU is not necessarily this container's.
_CFLAGS = ["-O3", "-mavx2", "-mfma", "-fopenmp", "-fPIC", "-shared", "-fno-fast-math"]

TARGETS = {
    "liboracle_topk.so": (["gcc"], [os.path.join(HERE, "c", "oracle_topk.c")], ["-lm"]),
    "libhnsw_baseline.so": (
        ["g++", "-std=c++17"],
        [os.path.join(HERE, "hnsw", "hnsw_baseline.cpp")],
        [],
    ),
}


def lib_path(name: str) -> str:
    sub = "c" if name == "liboracle_topk.so" else "hnsw"
    return os.path.join(HERE, sub, name)


def _stale(out: str, srcs: list[str]) -> bool:
    if not os.path.exists(out):
        return True
    mt = os.path.getmtime(out)
    return any(os.path.getmtime(s) > mt for s in srcs)


def build(name: str | None = None, force: bool = False, verbose: bool = False) -> None:
    names = [name] if name else list(TARGETS)
    for nm in names:
        cc, srcs, libs = TARGETS[nm]
        if not all(os.path.exists(s) for s in srcs):
            continue
        out = lib_path(nm)
        if not force and not _stale(out, srcs):
            continue
        cmd = cc + _CFLAGS + srcs + ["-o", out] + libs
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        subprocess.run(cmd, check=True)


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
