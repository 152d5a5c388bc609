"""Row-sharded corpus over the GPUs of one box.

The path shards naturally: the top-k over a union of disjoint row sets is the top-k of the
per-set top-k's.  Rank g holds the contiguous rows ``shard_bounds(N, G)[g]`` in its own
``DenseStore`` (``id_offset`` = first row, so "lower id wins" stays well defined across ranks);
queries are replicated; each rank produces a local top-k with the fp64 scores its ranking used;
ONE all-gather of ``[B, k] x (f64 score, i64 id)`` per batch (``B*k*16`` bytes per rank, over
NCCL / NVLink) is the only collective; the merge kernel (K5, ``cmw_merge_topk``) then reduces
``G*k -> k`` per query with the same (score desc, id asc) rule on every rank.

The reference has no counterpart (it talks to a single Chroma server:
rag_engine/storage/vector_store.py:34-42); this is the part that lets a corpus larger than one
GPU's 180 GB sit behind the same ``similarity_search_async`` seam.

``local_search`` / ``merge`` are injectable so that the host-side logic (bounds, offsets, gather
layout, padding of short shards) is testable under ``gloo`` on CPU with the oracle standing in
for the kernels; the defaults are the CUDA kernels and there is no CPU fallback.
"""
from __future__ import annotations

from typing import Callable


def shard_bounds(n_rows: int, world: int) -> list[tuple[int, int]]:
    """Contiguous row ranges [lo, hi) per rank: ceil(N / G) rows each, the last ones may be short
    or empty."""
    per = (n_rows + world - 1) // world if world > 0 else 0
    out = []
    for g in range(world):
        lo = min(n_rows, g * per)
        hi = min(n_rows, lo + per)
        out.append((lo, hi))
    return out


class ShardedSearcher:
    def __init__(self, store=None, group=None, local_search: Callable | None = None,
                 merge: Callable | None = None):
        import torch.distributed as dist

        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.store = store
        if local_search is None:
            if store is None:
                raise ValueError("ShardedSearcher needs a DenseStore or a local_search callable")

            def local_search(q, k, **kw):
                sc, ids, flags, s64 = store.search(q, k, return_scores64=True, **kw)
                return s64, ids, flags

        if merge is None:
            from .engine import merge_topk as merge
        self._local_search = local_search
        self._merge = merge

    def search(self, queries, k: int, **kw):
        """queries [B, dim] (replicated on every rank) -> (scores f32[B,k], ids i64[B,k], flags),
        identical on every rank."""
        import torch

        s64, ids, flags = self._local_search(queries, k, **kw)
        if self.world == 1:
            ms, mi, _ = self._merge(s64.unsqueeze(0), ids.unsqueeze(0), k)
            return ms, mi, flags
        b, kk = s64.shape
        # rank-major concatenation along dim 0 == a [world, B, k] stack
        g_s = torch.empty((self.world * b, kk), dtype=s64.dtype, device=s64.device)
        g_i = torch.empty((self.world * b, kk), dtype=ids.dtype, device=ids.device)
        self.dist.all_gather_into_tensor(g_s, s64.contiguous(), group=self.group)
        self.dist.all_gather_into_tensor(g_i, ids.contiguous(), group=self.group)
        g_s = g_s.view(self.world, b, kk)
        g_i = g_i.view(self.world, b, kk)
        if flags is not None:
            flags = flags.clone()
            self.dist.all_reduce(flags, op=self.dist.ReduceOp.MAX, group=self.group)
        ms, mi, _ = self._merge(g_s, g_i, k)
        return ms, mi, flags
