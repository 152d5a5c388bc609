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
        lib.oracle_topk_candidates_f32.restype = ctypes.c_int
        lib.oracle_topk_candidates_f32.argtypes = [
            ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
            ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
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


def exact_topk_candidates_c(corpus, queries, cand_ids, k, metric="cosine", nthreads=0):
    """Exact fp64 top-k among per-query candidate rows ``cand_ids`` i64[B, m] (negative = skip): the same
    arithmetic and order as :func:`exact_topk_c`.  Returns (ids int64[B,k], scores f32[B,k], scores f64[B,k])."""
    lib = _load()
    corpus = np.ascontiguousarray(corpus, dtype=np.float32)
    queries = np.ascontiguousarray(np.atleast_2d(queries), dtype=np.float32)
    cand = np.ascontiguousarray(cand_ids, dtype=np.int64)
    n, d = corpus.shape
    b, m = cand.shape
    assert queries.shape == (b, d)
    mt = {"cosine": 0, "ip": 1, 0: 0, 1: 1}[metric]
    ids = np.empty((b, k), np.int64)
    sc = np.empty((b, k), np.float64)
    rc = lib.oracle_topk_candidates_f32(corpus.ctypes.data, n, d, queries.ctypes.data, b, k, mt, cand.ctypes.data, m,
                                        ids.ctypes.data, sc.ctypes.data, nthreads)
    if rc != 0:
        raise RuntimeError(f"oracle_topk_candidates_f32 failed: {rc}")
    return ids, sc.astype(np.float32), sc


def exact_topk_prefiltered(corpus, queries, k, margin_rows=156, chunk=256, nthreads=0, check_margin=1e-4):
    """The oracle's exact top-k for MANY queries at full size: an fp32 sgemm prefilter (torch CPU, all host
    threads) keeps the best k + margin_rows rows per query, :func:`exact_topk_candidates_c` rescoring them in fp64
    in the oracle's order.  The prefilter is verified, not trusted: it is only valid if the weakest kept fp32 score
    lies more than ``check_margin`` (>> the fp32 dot-product error, ~1e-6 on unit vectors) below the k-th exact
    score -- asserted per query.  Cosine over unit-norm rows only (the bench corpus)."""
    import torch

    if nthreads:
        torch.set_num_threads(int(nthreads))
    c = torch.from_numpy(np.ascontiguousarray(corpus, dtype=np.float32))
    q = np.ascontiguousarray(queries, dtype=np.float32)
    b = q.shape[0]
    m = min(k + margin_rows, corpus.shape[0])
    ids = np.empty((b, k), np.int64)
    s32 = np.empty((b, k), np.float32)
    s64 = np.empty((b, k), np.float64)
    for lo in range(0, b, chunk):
        hi = min(b, lo + chunk)
        scores = torch.from_numpy(q[lo:hi]) @ c.T
        top_s, top_i = torch.topk(scores, m, dim=1, sorted=True)
        i_, s_, s64_ = exact_topk_candidates_c(corpus, q[lo:hi], top_i.numpy(), k, nthreads=nthreads)
        weakest = top_s[:, -1].numpy().astype(np.float64)
        if m < corpus.shape[0]:
            assert (s64_[:, k - 1] > weakest + check_margin).all(), "sgemm prefilter margin too small"
        ids[lo:hi], s32[lo:hi], s64[lo:hi] = i_, s_, s64_
    return ids, s32, s64
