"""ctypes wrapper over oracle/c/oracle_topk.c (TEST INFRASTRUCTURE ONLY)."""
from __future__ import annotations

import ctypes

import numpy as np

from . import build as _build

_lib = None


def _load():
    global _lib
    if _lib is None:
        _build.build("liboracle_topk.so")
        lib = ctypes.CDLL(_build.lib_path("liboracle_topk.so"))
        lib.oracle_topk_f32.restype = ctypes.c_int
        lib.oracle_topk_f32.argtypes = [
            ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
            ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
            ctypes.c_void_p, ctypes.c_int,
        ]
        lib.oracle_dot64_f32.restype = ctypes.c_double
        lib.oracle_dot64_f32.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        lib.oracle_num_threads.restype = ctypes.c_int
        _lib = lib
    return _lib


def num_threads() -> int:
    return int(_load().oracle_num_threads())


def exact_topk_c(corpus, queries, k, metric="cosine", live=None, id_offset=0, nthreads=0):
    """Same contract as oracle.topk.exact_topk, C/OpenMP speed.

    Returns (ids int64[B,k], scores f32[B,k], scores f64[B,k])."""
    lib = _load()
    corpus = np.ascontiguousarray(corpus, dtype=np.float32)
    queries = np.ascontiguousarray(np.atleast_2d(queries), dtype=np.float32)
    n, d = corpus.shape
    b = queries.shape[0]
    assert queries.shape[1] == d
    m = {"cosine": 0, "ip": 1, 0: 0, 1: 1}[metric]
    ids = np.empty((b, k), np.int64)
    sc = np.empty((b, k), np.float64)
    live_p = None
    if live is not None:
        live = np.ascontiguousarray(live, dtype=np.uint8)
        live_p = live.ctypes.data
    rc = lib.oracle_topk_f32(
        corpus.ctypes.data, n, d, queries.ctypes.data, b, k, m, live_p, id_offset,
        ids.ctypes.data, sc.ctypes.data, nthreads,
    )
    if rc != 0:
        raise RuntimeError(f"oracle_topk_f32 failed: {rc}")
    return ids, sc.astype(np.float32), sc
