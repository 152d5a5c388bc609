"""Exact cosine / inner-product top-k -- the ground truth (TEST INFRASTRUCTURE ONLY).

Restates what ``ChromaStore.similarity_search_async``
(``rag_engine/storage/vector_store.py:54-66``) asks its backend for: the ``k`` nearest
rows to ``query_embedding`` in the space declared at ``vector_store.py:48-51``
(``{"hnsw:space": "cosine"}``), best first.  Chroma/hnswlib answers approximately; the
oracle answers exactly, as ``BASELINE.json:north_star`` requires ("exact numpy/torch-CPU
cosine top-k is the ground truth ... ties broken by lower id").

Definition (the contract every GPU path is checked against):

    dot(q, c)  = sum_d float64(q[d]) * float64(c[d])            (fp64 accumulation)
    cosine     = dot(q, c) / sqrt(dot(q, q) * dot(c, c))         (0 if a norm is 0)
    ip         = dot(q, c)
    order      = score descending, then id ascending             (lower id wins ties)
    output     = ids int64[B, k], scores float32[B, k]; rows beyond the number of
                 live rows are (id = -1, score = -inf)           (k > N -> fewer results,
                                                                  as Chroma does [ext])

Candidate scores are re-evaluated with a row-position-independent summation so that
bit-identical rows get bit-identical fp64 scores (BLAS dgemm edge tiles may not).
"""
from __future__ import annotations

import numpy as np

COSINE = 0
IP = 1

_METRICS = {"cosine": COSINE, "ip": IP, COSINE: COSINE, IP: IP}


def _metric(metric) -> int:
    try:
        return _METRICS[metric]
    except KeyError:  # pragma: no cover - defensive
        raise ValueError(f"unknown metric {metric!r}") from None


def _canonical_dots(q64: np.ndarray, rows64: np.ndarray) -> np.ndarray:
    """fp64 dot of one query with a fresh contiguous [m, D] block (position independent)."""
    prod = np.ascontiguousarray(rows64 * q64[None, :])
    return prod.sum(axis=1)


def exact_scores(corpus: np.ndarray, queries: np.ndarray, metric="cosine") -> np.ndarray:
    """Full fp64 score matrix [B, N] (small cases only)."""
    m = _metric(metric)
    c64 = np.asarray(corpus, dtype=np.float64)
    q64 = np.asarray(queries, dtype=np.float64)
    s = q64 @ c64.T
    if m == COSINE:
        cn = np.sqrt(np.einsum("nd,nd->n", c64, c64))
        qn = np.sqrt(np.einsum("bd,bd->b", q64, q64))
        den = qn[:, None] * cn[None, :]
        with np.errstate(divide="ignore", invalid="ignore"):
            s = np.where(den > 0, s / den, 0.0)
    return s


def exact_topk(
    corpus: np.ndarray,
    queries: np.ndarray,
    k: int,
    metric="cosine",
    live: np.ndarray | None = None,
    id_offset: int = 0,
    block: int = 262144,
):
    """Exact top-k, (score desc, id asc).  Returns (ids int64[B,k], scores f32[B,k], scores64).

    ``live`` is an optional bool[N] mask (False = tombstoned row, never returned).
    ``id_offset`` is added to row numbers (row shards of a multi-GPU corpus).
    """
    m = _metric(metric)
    corpus = np.asarray(corpus)
    queries = np.atleast_2d(np.asarray(queries))
    n, d = corpus.shape
    b = queries.shape[0]
    assert queries.shape[1] == d
    q64 = queries.astype(np.float64)
    qn = np.sqrt(np.einsum("bd,bd->b", q64, q64))

    pad = 16  # extra candidates taken from the dgemm pass before canonical re-evaluation
    kk = min(n, k + pad)
    cand_ids = np.full((b, 0), -1, dtype=np.int64)
    cand_sc = np.full((b, 0), -np.inf, dtype=np.float64)
    for lo in range(0, n, block):
        hi = min(n, lo + block)
        c64 = corpus[lo:hi].astype(np.float64)
        s = q64 @ c64.T
        if m == COSINE:
            cn = np.sqrt(np.einsum("nd,nd->n", c64, c64))
            den = qn[:, None] * cn[None, :]
            with np.errstate(divide="ignore", invalid="ignore"):
                s = np.where(den > 0, s / den, 0.0)
        if live is not None:
            s = np.where(live[None, lo:hi], s, -np.inf)
        ids = np.broadcast_to(np.arange(lo, hi, dtype=np.int64)[None, :], s.shape)
        cand_ids = np.concatenate([cand_ids, ids], axis=1)
        cand_sc = np.concatenate([cand_sc, s], axis=1)
        if cand_sc.shape[1] > kk:
            # keep the kk best (ties at the cut are irrelevant: pad >> tie multiplicity
            # in every fixture; asserted below)
            part = np.argpartition(-cand_sc, kk - 1, axis=1)[:, :kk]
            cand_ids = np.take_along_axis(cand_ids, part, axis=1)
            cand_sc = np.take_along_axis(cand_sc, part, axis=1)

    out_ids = np.full((b, k), -1, dtype=np.int64)
    out_sc64 = np.full((b, k), -np.inf, dtype=np.float64)
    for i in range(b):
        ids = cand_ids[i]
        keep = np.isfinite(cand_sc[i])
        ids = ids[keep]
        if ids.size == 0:
            continue
        rows64 = corpus[ids].astype(np.float64)
        sc = _canonical_dots(q64[i], rows64)
        if m == COSINE:
            cn = np.sqrt(_canonical_dots_self(rows64))
            den = qn[i] * cn
            with np.errstate(divide="ignore", invalid="ignore"):
                sc = np.where(den > 0, sc / den, 0.0)
        order = np.lexsort((ids, -sc))
        take = order[:k]
        out_ids[i, : take.size] = ids[take] + id_offset
        out_sc64[i, : take.size] = sc[take]
    return out_ids, out_sc64.astype(np.float32), out_sc64


def _canonical_dots_self(rows64: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(rows64 * rows64).sum(axis=1)


def merge_topk(ids: np.ndarray, scores: np.ndarray, k: int):
    """Merge G candidate lists per query: ids/scores [G, B, k'] -> top-k by (score desc, id asc).

    Restates the cross-shard exchange of SURVEY.md §8(e): the top-k of a union of
    disjoint row sets is the top-k of the per-set top-k's.  Entries with id < 0 are padding.
    """
    ids = np.asarray(ids)
    scores = np.asarray(scores)
    g, b, kp = ids.shape
    flat_ids = np.transpose(ids, (1, 0, 2)).reshape(b, g * kp)
    flat_sc = np.transpose(scores, (1, 0, 2)).reshape(b, g * kp).astype(np.float64)
    flat_sc = np.where(flat_ids < 0, -np.inf, flat_sc)
    out_ids = np.full((b, k), -1, dtype=np.int64)
    out_sc = np.full((b, k), -np.inf, dtype=np.float32)
    for i in range(b):
        key_ids = np.where(flat_ids[i] < 0, np.iinfo(np.int64).max, flat_ids[i])
        order = np.lexsort((key_ids, -flat_sc[i]))[:k]
        valid = flat_ids[i][order] >= 0
        order = order[valid]
        out_ids[i, : order.size] = flat_ids[i][order]
        out_sc[i, : order.size] = flat_sc[i][order].astype(np.float32)
    return out_ids, out_sc
