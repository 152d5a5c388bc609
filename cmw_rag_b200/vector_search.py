"""Vector-search seam -- mirror of rag_engine/retrieval/vector_search.py:8-10 of the reference.

The reference's own function works unchanged with a ``B200Store`` (it is duck-typed); this copy
exists so callers that do not have the reference importable get the same entry point.
"""
from __future__ import annotations

from typing import Any, List


async def top_k_search_async(store, embedding: List[float], k: int) -> List[Any]:
    """Async: top-k results from the store for one query embedding, best first."""
    return await store.similarity_search_async(query_embedding=embedding, k=k)
