"""ctypes wrapper over oracle/hnsw/hnsw_baseline.cpp -- the "Chroma HNSW" CPU baseline
(hnswlib-equivalent re-implementation; chromadb 1.3.0 defaults assumed, not verifiable offline).
TEST / BENCHMARK INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes

import numpy as np

from . import build as _build

_lib = None

# Chroma 1.x collection defaults (only hnsw:space is set by the reference:
# rag_engine/storage/vector_store.py:48-51)
CHROMA_M = 16
CHROMA_EF_CONSTRUCTION = 100
CHROMA_EF_SEARCH = 100


def _load():
    global _lib
    if _lib is None:
        _build.build("libhnsw_baseline.so")
        lib = ctypes.CDLL(_build.lib_path("libhnsw_baseline.so"))
        lib.hnsw_create.restype = ctypes.c_void_p
        lib.hnsw_create.argtypes = [ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_uint64]
        lib.hnsw_free.argtypes = [ctypes.c_void_p]
        lib.hnsw_add.restype = ctypes.c_int64
        lib.hnsw_add.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int]
        lib.hnsw_search.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                    ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        lib.hnsw_threads.restype = ctypes.c_int
        _lib = lib
    return _lib


def num_threads() -> int:
    return int(_load().hnsw_threads())


class HnswIndex:
    def __init__(self, dim: int, max_elements: int, M: int = CHROMA_M,
                 ef_construction: int = CHROMA_EF_CONSTRUCTION, seed: int = 100):
        self._lib = _load()
        self.dim = dim
        self._h = self._lib.hnsw_create(dim, max_elements, M, ef_construction, seed)
        if not self._h:
            raise MemoryError("hnsw_create failed")

    def add(self, rows: np.ndarray, nthreads: int = 0) -> int:
        rows = np.ascontiguousarray(rows, dtype=np.float32)
        assert rows.ndim == 2 and rows.shape[1] == self.dim
        n = int(self._lib.hnsw_add(self._h, rows.ctypes.data, rows.shape[0], nthreads))
        if n < 0:
            raise ValueError("index full")
        return n

    def search(self, queries: np.ndarray, k: int, ef_search: int = CHROMA_EF_SEARCH, nthreads: int = 0):
        q = np.ascontiguousarray(np.atleast_2d(queries), dtype=np.float32)
        ids = np.empty((q.shape[0], k), np.int64)
        dist = np.empty((q.shape[0], k), np.float32)
        self._lib.hnsw_search(self._h, q.ctypes.data, q.shape[0], k, ef_search, ids.ctypes.data,
                              dist.ctypes.data, nthreads)
        return ids, dist

    def close(self):
        if self._h:
            self._lib.hnsw_free(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass
