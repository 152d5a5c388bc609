"""cmw_rag_b200 -- B200-native dense-retrieval hot path of cmw-rag.

Host layer (Python, mirrors the reference's store / seam interface) over ``libcmwdense.so``
(C ABI, hand-written sm_100a CUDA kernels).  Importing this package does not need a GPU; every
compute call does, and fails loudly without one -- there is no CPU fallback.
"""
from . import _native as native  # noqa: F401
from .kbid import extract_numeric_kbid, group_key  # noqa: F401
from .vector_search import top_k_search_async  # noqa: F401


def __getattr__(name):  # lazy: engine/store import numpy-only modules but keep import light
    if name in ("DenseStore", "merge_topk", "MultiVectorResult"):
        from . import engine

        return getattr(engine, name)
    if name in ("B200Store", "RetrievedDoc"):
        from . import store

        return getattr(store, name)
    if name in ("ShardedSearcher",):
        from . import sharded

        return getattr(sharded, name)
    if name in ("SearchBatcher",):
        from . import batcher

        return getattr(batcher, name)
    if name in ("CollectionRegistry",):
        from . import registry

        return getattr(registry, name)
    raise AttributeError(name)
