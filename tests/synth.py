"""Synthetic FRIDA-shaped data, as specified in SURVEY.md §8(d).  numpy only (host side).

Corpus rows iid N(0,1) then L2-normalised in fp32 (FRIDA emits normalised 1536-d vectors:
rag_engine/retrieval/embedder.py:143-148 of the reference, dim at config/models.yaml:8-11).
Queries: 75 % planted needles ``normalise(C[j] + 0.75 g)`` (cos ~ 0.8 to row j, so
top-1 must be j at any N), 25 % pure random unit vectors.
"""
from __future__ import annotations

import numpy as np

CORPUS_SEED = 20261018
QUERY_SEED = 7
KB_SEED = 11


def _normalise(x: np.ndarray) -> np.ndarray:
    n = np.sqrt((x.astype(np.float64) ** 2).sum(axis=1, keepdims=True))
    return (x / np.maximum(n, 1e-30)).astype(np.float32)


def make_corpus(n: int, d: int = 1536, seed: int = CORPUS_SEED, ties: bool = True) -> np.ndarray:
    rng = np.random.Generator(np.random.PCG64(seed))
    c = np.empty((n, d), np.float32)
    step = 65536
    for lo in range(0, n, step):
        hi = min(n, lo + step)
        c[lo:hi] = _normalise(rng.standard_normal((hi - lo, d), dtype=np.float32))
    if ties and n >= 64:
        # tie fixtures: row 17 duplicated at N/2 and N-1 (lower id must win)
        c[n // 2] = c[17]
        c[n - 1] = c[17]
    return c


def make_queries(corpus: np.ndarray, b: int, seed: int = QUERY_SEED, tie_probe: bool = True):
    """Returns (Q f32[b,d], needle int64[b]) -- needle[i] = planted row or -1."""
    n, d = corpus.shape
    rng = np.random.Generator(np.random.PCG64(seed))
    needle = rng.integers(0, n, size=b)
    g = rng.standard_normal((b, d), dtype=np.float32)
    g = _normalise(g)
    q = corpus[needle] + 0.75 * g
    rnd = rng.random(b) < 0.25
    q[rnd] = g[rnd]
    needle = np.where(rnd, -1, needle).astype(np.int64)
    if tie_probe and b >= 4 and n >= 64:
        # query 1 is planted on the duplicated row: expects ids 17, N/2, N-1 in that order
        q[1] = corpus[17] + 0.75 * g[1]
        needle[1] = 17
    return _normalise(q), needle


def make_clustered_corpus(n: int, d: int = 1536, n_centroids: int = 4096, seed: int = CORPUS_SEED + 1):
    rng = np.random.Generator(np.random.PCG64(seed))
    cent = _normalise(rng.standard_normal((n_centroids, d), dtype=np.float32))
    assign = rng.integers(0, n_centroids, size=n)
    c = np.empty((n, d), np.float32)
    step = 65536
    for lo in range(0, n, step):
        hi = min(n, lo + step)
        noise = _normalise(rng.standard_normal((hi - lo, d), dtype=np.float32))
        c[lo:hi] = _normalise(cent[assign[lo:hi]] + 0.35 * noise)
    return c


def make_kbids(n: int, seed: int = KB_SEED):
    """Articles of geometric length (mean 8 chunks) laid out contiguously.

    Returns (kb_strings list[str] len n, article_no int64[n]).  kbId strings are
    ``str(1000+g)``; 1 % carry a ``-toc`` suffix (exercises extract_numeric_kbid,
    metadata_utils.py:31), 0.1 % are empty (dropped per retriever.py:238).
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    lens = rng.geometric(1.0 / 8.0, size=max(16, n // 4 + 16))
    art = np.repeat(np.arange(lens.size), lens)[:n]
    assert art.size == n
    u = rng.random(n)
    kb = []
    for i in range(n):
        s = str(1000 + int(art[i]))
        if u[i] < 0.001:
            s = ""
        elif u[i] < 0.011:
            s += "-toc"
        kb.append(s)
    return kb, art.astype(np.int64)
