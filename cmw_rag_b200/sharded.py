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


class PeerExchange:
    """The NVLink peer-memory alternative to all-gather + merge (cmw_exchange_merge, exchange.cu).

    Every rank allocates one peer buffer inside libcmwdense.so, the 64-byte cudaIpc handles travel
    through ``torch.distributed`` once at construction, and from then on an exchange is two kernels
    of ours on the caller's stream: peer stores of the local candidates into every rank's buffer +
    epoch flags, then a flag-wait + merge -- no NCCL call on the data path."""

    def __init__(self, group=None, device: int = 0, max_batch: int = 4096, max_k: int = 128):
        import ctypes

        import torch.distributed as dist

        from . import _native as N

        self._N = N
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.device = int(device)
        self.max_batch, self.max_k = int(max_batch), int(max_k)
        lib = N.lib()
        nbytes = int(lib.cmw_peer_buffer_bytes(self.world, self.max_batch, self.max_k))
        own = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(64)
        N.check(lib.cmw_peer_alloc(self.device, nbytes, ctypes.byref(own), handle), "cmw_peer_alloc")
        handles: list = [None] * self.world
        dist.all_gather_object(handles, bytes(handle.raw), group=group)
        self._own = own
        self._opened = []
        ptrs = []
        for g in range(self.world):
            if g == self.rank:
                ptrs.append(own.value)
                continue
            p = ctypes.c_void_p()
            N.check(lib.cmw_peer_open(self.device, ctypes.create_string_buffer(handles[g], 64), ctypes.byref(p)),
                    "cmw_peer_open")
            self._opened.append(p)
            ptrs.append(p.value)
        self._ptrs = (ctypes.c_void_p * self.world)(*ptrs)
        self.epoch = 0
        dist.barrier(group=group)

    def exchange_merge(self, s64, ids, k_out: int):
        """s64 f64[B,k], ids i64[B,k] (CUDA, this rank's candidates) -> merged (scores f32, ids, scores64)."""
        import torch

        assert s64.is_cuda and s64.dtype == torch.float64 and ids.dtype == torch.int64 and s64.shape == ids.shape
        s64, ids = s64.contiguous(), ids.contiguous()
        b, k = s64.shape
        if b > self.max_batch or k > self.max_k:
            raise ValueError(f"exchange sized for batch {self.max_batch} x k {self.max_k}, got {b} x {k}")
        dev = s64.device
        out_s = torch.empty((b, k_out), dtype=torch.float32, device=dev)
        out_i = torch.empty((b, k_out), dtype=torch.int64, device=dev)
        out_s64 = torch.empty((b, k_out), dtype=torch.float64, device=dev)
        self.epoch += 1
        stream = torch.cuda.current_stream(dev).cuda_stream
        self._N.check(
            self._N.lib().cmw_exchange_merge(self._ptrs, self.world, self.rank, self.max_batch, self.max_k, b, k,
                                             k_out, self.epoch & 0xffffffff or 2, s64.data_ptr(), ids.data_ptr(),
                                             out_s.data_ptr(), out_i.data_ptr(), out_s64.data_ptr(), stream),
            "cmw_exchange_merge")
        return out_s, out_i, out_s64

    def close(self):
        import torch
        import torch.distributed as dist

        if self._own is None:
            return
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)  # nobody may still be writing into a buffer that is about to go
        lib = self._N.lib()
        for p in self._opened:
            lib.cmw_peer_close(p)
        lib.cmw_peer_free(self._own)
        self._own, self._opened = None, []


class ShardedSearcher:
    def __init__(self, store=None, group=None, local_search: Callable | None = None,
                 merge: Callable | None = None, exchange: PeerExchange | None = None):
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
        self._exchange = exchange

    def search(self, queries, k: int, **kw):
        """queries [B, dim] (replicated on every rank) -> (scores f32[B,k], ids i64[B,k], flags),
        identical on every rank."""
        import torch

        s64, ids, flags = self._local_search(queries, k, **kw)
        if self.world == 1:
            ms, mi, _ = self._merge(s64.unsqueeze(0), ids.unsqueeze(0), k)
            return ms, mi, flags
        if self._exchange is not None:
            # fused path: peer stores over NVLink + flag-wait + merge, no collective call
            if flags is not None:
                flags = flags.clone()
                self.dist.all_reduce(flags, op=self.dist.ReduceOp.MAX, group=self.group)
            ms, mi, _ = self._exchange.exchange_merge(s64, ids, k)
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
