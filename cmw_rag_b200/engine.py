"""DenseStore -- the Python handle on one HBM-resident corpus shard (``cmw_store`` in the C ABI).

PyTorch is only the hand-off: tensors own query / output / workspace memory, and their
``data_ptr()`` plus the current CUDA stream are what crosses into ``libcmwdense.so``.  Every
compute method ends in a kernel of that library; nothing here computes a score or a top-k in
Python, numpy or torch.

Replaces, on the reference side, the Chroma collection behind
``rag_engine/storage/vector_store.py:44-66`` (create + query) for whole batches of queries.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass

import numpy as np

from . import _native as N


@dataclass
class MultiVectorResult:
    """Outputs of K4 for Q long queries (arrays padded to P = pre-rerank cap or S*k).

    Mirrors the intermediate values of ``RAGRetriever.retrieve_async``
    (rag_engine/retrieval/retriever.py:185-242,307): ``cand_*`` is the deduplicated candidate
    list in first-seen order, ``grp_*`` the kbId groups in first-appearance order and
    ``grp_order`` their stable score-descending order.
    """

    cand_ids: "object"
    cand_scores: "object"
    cand_best: "object"
    cand_n: "object"
    cand_grp: "object"
    grp_gid: "object"
    grp_max: "object"
    grp_cnt: "object"
    grp_first: "object"
    grp_order: "object"
    grp_n: "object"

    def cpu(self) -> "MultiVectorResult":
        return MultiVectorResult(**{k: v.cpu() for k, v in self.__dict__.items()})


def _torch():
    import torch

    return torch


class DenseStore:
    """One corpus shard on one B200.

    ``dim``: embedding width (FRIDA: 1536 -- rag_engine/config/models.yaml:8-11 of the reference).
    ``capacity``: rows reserved in HBM up front (pointers never move, TMA descriptors stay valid).
    ``id_offset``: added to row numbers in results (row shards of a multi-GPU corpus).
    ``bf16``: keep 16-bit tiles of the normalised rows (the tensor-core filter's operand); ``tiles16`` picks
    their format -- ``"f16"`` (default: 8x smaller rounding residual, so the rigorous exactness certificate
    costs nothing) or ``"bf16"`` (BASELINE.json's literal format).
    """

    def __init__(self, dim: int, capacity: int, device: int = 0, f32: bool = True, bf16: bool = True,
                 id_offset: int = 0, tiles16: str = "f16"):
        lib = N.lib()
        if tiles16 not in ("f16", "bf16"):
            raise ValueError(f"tiles16 must be 'f16' or 'bf16', got {tiles16!r}")
        t16 = (N.STORE_F16 if tiles16 == "f16" else N.STORE_BF16) if bf16 else 0
        flags = (N.STORE_F32 if f32 else 0) | t16
        self.tiles16 = tiles16 if bf16 else None
        handle = ctypes.c_void_p()
        N.check(lib.cmw_store_create(int(device), int(dim), int(capacity), flags, int(id_offset),
                                     ctypes.byref(handle)), "cmw_store_create")
        self._h = handle
        self.dim = int(dim)
        self.device = int(device)
        self.capacity = int(capacity)
        self.id_offset = int(id_offset)
        self._ws = {}
        self._tickets = {}

    # -- lifecycle ---------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None):
            N.lib().cmw_store_destroy(self._h)
            self._h = None
            self._ws = {}

    def __del__(self):  # pragma: no cover - best effort
        try:
            self.close()
        except Exception:
            pass

    def info(self) -> dict:
        st = N.StoreInfo()
        N.check(N.lib().cmw_store_get_info(self._h, ctypes.byref(st)), "cmw_store_get_info")
        d = {name: int(getattr(st, name)) for name, _ in N.StoreInfo._fields_}
        d["gemm_ready"] = bool(d["flags"] & 0x100)
        return d

    @property
    def rows(self) -> int:
        return self.info()["rows"]

    @property
    def live_rows(self) -> int:
        return self.info()["live_rows"]

    # -- ingest (K0) ---------------------------------------------------------------------------
    def append(self, rows, kb_gid=None) -> None:
        """Append fp32 rows [n, dim] (numpy array -> staged H2D; CUDA tensor -> direct)."""
        torch = _torch()
        lib = N.lib()
        if isinstance(rows, torch.Tensor) and rows.is_cuda:
            rows = rows.contiguous().to(torch.float32)
            assert rows.dim() == 2 and rows.shape[1] == self.dim
            gid_ptr = None
            if kb_gid is not None:
                kb_gid = torch.as_tensor(kb_gid, dtype=torch.int32, device=rows.device).contiguous()
                assert kb_gid.numel() == rows.shape[0]
                gid_ptr = kb_gid.data_ptr()
            stream = torch.cuda.current_stream(rows.device).cuda_stream
            N.check(lib.cmw_store_append_f32(self._h, rows.data_ptr(), gid_ptr, rows.shape[0], stream),
                    "cmw_store_append_f32")
            # the kernel reads `rows` asynchronously; keep the caller's tensor alive until it ran
            torch.cuda.current_stream(rows.device).synchronize()
            return
        arr = np.ascontiguousarray(rows.cpu().numpy() if isinstance(rows, torch.Tensor) else rows,
                                   dtype=np.float32)
        assert arr.ndim == 2 and arr.shape[1] == self.dim, f"expected [n, {self.dim}], got {arr.shape}"
        gid_ptr = None
        if kb_gid is not None:
            gid = np.ascontiguousarray(kb_gid, dtype=np.int32)
            assert gid.shape[0] == arr.shape[0]
            gid_ptr = gid.ctypes.data
        N.check(lib.cmw_store_append_host_f32(self._h, arr.ctypes.data, gid_ptr, arr.shape[0]),
                "cmw_store_append_host_f32")

    def copy_rows_from(self, src: "DenseStore", rows=None, row0: int = 0, n: int | None = None) -> None:
        """Append rows of ``src`` (same GPU) to this store without leaving the device: ``rows`` = LOCAL row numbers
        of ``src`` (array / tensor), or the contiguous range [row0, row0 + n).  K0 rebuilds tiles and norms."""
        torch = _torch()
        dev = torch.device(f"cuda:{self.device}")
        stream = torch.cuda.current_stream(dev).cuda_stream
        if rows is not None:
            idx = torch.as_tensor(np.asarray(rows, dtype=np.int64) if not isinstance(rows, torch.Tensor) else rows,
                                  dtype=torch.int64, device=dev).contiguous()
            if idx.numel():
                N.check(N.lib().cmw_store_copy_rows(self._h, src._h, idx.data_ptr(), 0, idx.numel(), stream),
                        "cmw_store_copy_rows")
                torch.cuda.current_stream(dev).synchronize()  # the index tensor must outlive the kernel
            return
        n = src.rows - row0 if n is None else n
        if n:
            N.check(N.lib().cmw_store_copy_rows(self._h, src._h, None, int(row0), int(n), stream),
                    "cmw_store_copy_rows")
            torch.cuda.current_stream(dev).synchronize()  # the caller may close `src` right away

    def read_rows(self, row0: int, n: int, rows: bool = True):
        """Read LOCAL rows back: (f32 [n, dim] or None, kb_gid i32[n], live bool[n])."""
        out = np.empty((n, self.dim), np.float32) if rows else None
        gid = np.empty((n,), np.int32)
        live = np.empty((n,), np.uint8)
        if n:
            N.check(N.lib().cmw_store_read_rows_f32(self._h, int(row0), int(n),
                                                    out.ctypes.data if rows else None, gid.ctypes.data,
                                                    live.ctypes.data), "cmw_store_read_rows_f32")
        return out, gid, live.astype(bool)

    def tombstone(self, rows) -> None:
        """Mark LOCAL row numbers dead (never returned again)."""
        arr = np.ascontiguousarray(rows, dtype=np.int64)
        if arr.size:
            N.check(N.lib().cmw_store_tombstone_host(self._h, arr.ctypes.data, arr.size),
                    "cmw_store_tombstone_host")

    # -- search -----------------------------------------------------------------------------------
    @staticmethod
    def _mode(mode, algo) -> int:
        return N.MODES[mode] | N.ALGOS[algo]

    def _workspace(self, batch: int, k: int, mode: int, slot: int = 0):
        """One workspace per slot: searches that may overlap (different streams) need different slots."""
        torch = _torch()
        need = int(N.lib().cmw_search_workspace_bytes(self._h, batch, k, mode))
        ws = self._ws.get(slot)
        if ws is None or ws.numel() < need:
            ws = torch.empty(need, dtype=torch.uint8, device=f"cuda:{self.device}")
            self._ws[slot] = ws
        return ws

    def search(self, queries, k: int, metric="cosine", mode="f32", algo=None, return_scores64=False, ws_slot: int = 0):
        """Batched top-k on the device.  ``queries``: CUDA fp32 tensor [B, dim].

        Returns (scores f32[B,k], ids i64[B,k], flags i32[B]) as CUDA tensors, stream-ordered on
        the current stream (plus scores64 f64[B,k] when asked: needed for an exact shard merge).
        """
        torch = _torch()
        assert queries.is_cuda and queries.dtype == torch.float32 and queries.dim() == 2
        assert queries.shape[1] == self.dim
        q = queries.contiguous()
        b = q.shape[0]
        dev = q.device
        m = self._mode(mode, algo)
        scores = torch.empty((b, k), dtype=torch.float32, device=dev)
        ids = torch.empty((b, k), dtype=torch.int64, device=dev)
        flags = torch.zeros((b,), dtype=torch.int32, device=dev)
        s64 = torch.empty((b, k), dtype=torch.float64, device=dev) if return_scores64 else None
        if b == 0:
            return (scores, ids, flags, s64) if return_scores64 else (scores, ids, flags)
        ws = self._workspace(b, k, m, ws_slot)
        stream = torch.cuda.current_stream(dev).cuda_stream
        N.check(
            N.lib().cmw_search(self._h, q.data_ptr(), b, k, N.METRICS[metric], m, scores.data_ptr(),
                               ids.data_ptr(), s64.data_ptr() if s64 is not None else None,
                               flags.data_ptr(), ws.data_ptr(), ws.numel(), stream),
            "cmw_search",
        )
        return (scores, ids, flags, s64) if return_scores64 else (scores, ids, flags)

    # -- row-sharded search: the two halves of a search around the cross-shard exchange (sharded.py) -------
    def search_filter(self, queries, k: int, metric="cosine", mode="f32", algo=None, ws_slot: int = 0):
        """First half (prep + filter + compaction).  Returns the best k FILTER scores f32[B,k] of this
        shard; the candidate pools stay in the workspace of ``ws_slot`` for :meth:`search_finish`."""
        torch = _torch()
        assert queries.is_cuda and queries.dtype == torch.float32 and queries.dim() == 2
        q = queries.contiguous()
        b = q.shape[0]
        m = self._mode(mode, algo)
        ftop = torch.empty((b, k), dtype=torch.float32, device=q.device)
        if b:
            ws = self._workspace(b, k, m, ws_slot)
            N.check(N.lib().cmw_search_filter(self._h, q.data_ptr(), b, k, N.METRICS[metric], m, ftop.data_ptr(),
                                              ws.data_ptr(), ws.numel(),
                                              torch.cuda.current_stream(q.device).cuda_stream), "cmw_search_filter")
        return ftop

    def search_finish(self, queries, k: int, global_kth=None, metric="cosine", mode="f32", algo=None,
                      ws_slot: int = 0, block=None):
        """Second half: rescoring (restricted by the cross-shard k-th filter score ``global_kth`` f32[B] when
        given) + selection into one packed uint8 block (cmw_shard_block_bytes) for the all-gather."""
        torch = _torch()
        q = queries.contiguous()
        b = q.shape[0]
        m = self._mode(mode, algo)
        nbytes = int(N.lib().cmw_shard_block_bytes(max(b, 1), k))
        if block is None:
            block = torch.empty(nbytes, dtype=torch.uint8, device=q.device)
        assert block.numel() >= nbytes and block.is_cuda
        if b:
            ws = self._workspace(b, k, m, ws_slot)
            N.check(N.lib().cmw_search_finish(self._h, q.data_ptr(), b, k, N.METRICS[metric], m,
                                              global_kth.data_ptr() if global_kth is not None else None,
                                              block.data_ptr(), ws.data_ptr(), ws.numel(),
                                              torch.cuda.current_stream(q.device).cuda_stream), "cmw_search_finish")
        return block

    def search_host(self, queries, k: int, metric="cosine", mode="f32", algo=None, out=None):
        """End-to-end form with HOST buffers (numpy in, numpy out): H2D, kernels, D2H, synchronised.
        This is the call that stands in for one HTTP round trip to Chroma.  Pageable arrays go through
        the store's pinned staging buffer; page-locked ones (``pinned_empty``) are used for DMA
        directly.  ``out`` = optional preallocated (scores f32[B,k], ids i64[B,k], flags i32[B])."""
        q = np.ascontiguousarray(np.atleast_2d(queries), dtype=np.float32)
        assert q.shape[1] == self.dim
        b = q.shape[0]
        if out is not None:
            scores, ids, flags = out
            assert scores.shape == (b, k) and scores.dtype == np.float32 and scores.flags.c_contiguous
            assert ids.shape == (b, k) and ids.dtype == np.int64 and ids.flags.c_contiguous
            assert flags.shape == (b,) and flags.dtype == np.int32
        else:
            scores = np.empty((b, k), np.float32)
            ids = np.empty((b, k), np.int64)
            flags = np.zeros((b,), np.int32)
        if b:
            N.check(
                N.lib().cmw_search_host(self._h, q.ctypes.data, b, k, N.METRICS[metric],
                                        self._mode(mode, algo), scores.ctypes.data, ids.ctypes.data,
                                        flags.ctypes.data),
                "cmw_search_host",
            )
        return scores, ids, flags

    def search_host_submit(self, queries, k: int, metric="cosine", mode="f32", algo=None, out=None) -> int:
        """Pipelined form of :meth:`search_host`: enqueue H2D -> search -> D2H and return a ticket at once.
        Up to ``N.HOST_SLOTS`` tickets may be in flight; the copies of one overlap the kernels of the
        others.  :meth:`search_host_wait` returns the (scores, ids, flags) arrays of a ticket."""
        q = np.ascontiguousarray(np.atleast_2d(queries), dtype=np.float32)
        assert q.shape[1] == self.dim and q.shape[0] > 0
        b = q.shape[0]
        if out is not None:
            scores, ids, flags = out
            assert scores.shape == (b, k) and scores.dtype == np.float32 and scores.flags.c_contiguous
            assert ids.shape == (b, k) and ids.dtype == np.int64 and ids.flags.c_contiguous
            assert flags.shape == (b,) and flags.dtype == np.int32
        else:
            scores = np.empty((b, k), np.float32)
            ids = np.empty((b, k), np.int64)
            flags = np.zeros((b,), np.int32)
        ticket = ctypes.c_int(-1)
        N.check(
            N.lib().cmw_search_host_submit(self._h, q.ctypes.data, b, k, N.METRICS[metric], self._mode(mode, algo),
                                           scores.ctypes.data, ids.ctypes.data, flags.ctypes.data,
                                           ctypes.byref(ticket)),
            "cmw_search_host_submit",
        )
        # the library reads `q` and writes the outputs until the wait: keep them alive with the ticket
        self._tickets[ticket.value] = (q, scores, ids, flags)
        return ticket.value

    def search_host_wait(self, ticket: int):
        q, scores, ids, flags = self._tickets.pop(ticket)
        N.check(N.lib().cmw_search_host_wait(self._h, int(ticket)), "cmw_search_host_wait")
        return scores, ids, flags

    # -- multi-vector reduction (K4) -----------------------------------------------------------------
    def multivector(self, ids, scores, prl: int = 0, limit: int = 0, kb_gid=None, kb_id_offset=None):
        """ids i64[Q,S,k] / scores f32[Q,S,k] (CUDA) -> MultiVectorResult (CUDA tensors).

        ``kb_gid`` defaults to this store's own table; pass a full-corpus table (and its id offset)
        when the ids come from a cross-shard merge."""
        torch = _torch()
        assert ids.is_cuda and ids.dtype == torch.int64 and ids.dim() == 3
        assert scores.shape == ids.shape and scores.dtype == torch.float32
        ids = ids.contiguous()
        scores = scores.contiguous()
        qn, s, k = ids.shape
        n = s * k
        p = prl if 0 < prl < n else n
        dev = ids.device
        if kb_gid is None:
            kb_ptr = N.lib().cmw_store_kb_gid_dev(self._h)
            kb_rows = self.info()["rows"]
            kb_off = self.id_offset
        else:
            assert kb_gid.is_cuda and kb_gid.dtype == torch.int32
            kb_gid = kb_gid.contiguous()
            kb_ptr, kb_rows = kb_gid.data_ptr(), kb_gid.numel()
            kb_off = int(kb_id_offset or 0)
        out = MultiVectorResult(
            cand_ids=torch.empty((qn, p), dtype=torch.int64, device=dev),
            cand_scores=torch.empty((qn, p), dtype=torch.float32, device=dev),
            cand_best=torch.empty((qn, p), dtype=torch.float32, device=dev),
            cand_n=torch.zeros((qn,), dtype=torch.int32, device=dev),
            cand_grp=torch.empty((qn, p), dtype=torch.int32, device=dev),
            grp_gid=torch.empty((qn, p), dtype=torch.int32, device=dev),
            grp_max=torch.empty((qn, p), dtype=torch.float32, device=dev),
            grp_cnt=torch.empty((qn, p), dtype=torch.int32, device=dev),
            grp_first=torch.empty((qn, p), dtype=torch.int32, device=dev),
            grp_order=torch.empty((qn, p), dtype=torch.int32, device=dev),
            grp_n=torch.zeros((qn,), dtype=torch.int32, device=dev),
        )
        if qn == 0:
            return out
        stream = torch.cuda.current_stream(dev).cuda_stream
        N.check(
            N.lib().cmw_multivector(
                kb_ptr, kb_rows, kb_off, ids.data_ptr(), scores.data_ptr(), qn, s, k, int(prl), int(limit),
                out.cand_ids.data_ptr(), out.cand_scores.data_ptr(), out.cand_best.data_ptr(),
                out.cand_n.data_ptr(), out.cand_grp.data_ptr(), out.grp_gid.data_ptr(),
                out.grp_max.data_ptr(), out.grp_cnt.data_ptr(), out.grp_first.data_ptr(),
                out.grp_order.data_ptr(), out.grp_n.data_ptr(), stream),
            "cmw_multivector",
        )
        return out

    def search_multivector(self, segment_queries, k: int, prl: int = 0, limit: int = 0, metric="cosine",
                           mode="f32", algo=None):
        """[Q, S, dim] segment embeddings -> per-segment top-k (one batched launch instead of the
        reference's S awaits, retriever.py:179-182) -> union / dedup / cap / kbId groups."""
        torch = _torch()
        qn, s, d = segment_queries.shape
        flat = segment_queries.reshape(qn * s, d)
        scores, ids, flags = self.search(flat, k, metric=metric, mode=mode, algo=algo)
        res = self.multivector(ids.view(qn, s, k), scores.view(qn, s, k), prl=prl, limit=limit)
        return res, scores.view(qn, s, k), ids.view(qn, s, k), flags.view(qn, s)


def pinned_empty(shape, dtype) -> np.ndarray:
    """A page-locked numpy array (backed by a torch pinned tensor kept alive by the array)."""
    torch = _torch()
    t = torch.empty(tuple(shape), dtype=getattr(torch, np.dtype(dtype).name), pin_memory=True)
    return t.numpy()


def merge_topk(scores64, ids, k_out: int):
    """K5: scores f64[G,B,k_in], ids i64[G,B,k_in] (CUDA) -> (scores f32[B,k_out], ids, scores64)."""
    torch = _torch()
    assert scores64.is_cuda and scores64.dtype == torch.float64 and ids.dtype == torch.int64
    scores64 = scores64.contiguous()
    ids = ids.contiguous()
    g, b, k_in = ids.shape
    dev = ids.device
    out_s = torch.empty((b, k_out), dtype=torch.float32, device=dev)
    out_i = torch.empty((b, k_out), dtype=torch.int64, device=dev)
    out_s64 = torch.empty((b, k_out), dtype=torch.float64, device=dev)
    if b:
        stream = torch.cuda.current_stream(dev).cuda_stream
        N.check(
            N.lib().cmw_merge_topk(scores64.data_ptr(), ids.data_ptr(), g, b, k_in, k_out,
                                   out_s.data_ptr(), out_i.data_ptr(), out_s64.data_ptr(), stream),
            "cmw_merge_topk",
        )
    return out_s, out_i, out_s64


class DevicePtr:
    """A raw device pointer to `nbytes` of gathered data that lives outside torch (in a peer buffer of
    libcmwdense.so: PeerGather); `status` = the exchange's i32[1] status tensor."""

    __slots__ = ("ptr", "nbytes", "device", "status")

    def __init__(self, ptr: int, nbytes: int, device, status=None):
        self.ptr, self.nbytes, self.device, self.status = int(ptr), int(nbytes), device, status


def shard_kth(filter_topk_gathered, k: int, world: int | None = None, batch: int | None = None):
    """f32[G,B,k] gathered per-shard filter scores (a tensor, or a DevicePtr with `world` and `batch`) -> f32[B]:
    the k-th best over all shards."""
    torch = _torch()
    if isinstance(filter_topk_gathered, DevicePtr):
        g, b, dev, ptr = int(world), int(batch), filter_topk_gathered.device, filter_topk_gathered.ptr
        assert filter_topk_gathered.nbytes >= g * b * k * 4
    else:
        t = filter_topk_gathered.contiguous()
        g, b, kk = t.shape
        assert kk == k and t.dtype == torch.float32 and t.is_cuda
        dev, ptr = t.device, t.data_ptr()
    out = torch.empty((b,), dtype=torch.float32, device=dev)
    if b:
        N.check(N.lib().cmw_shard_kth(ptr, g, b, k, out.data_ptr(), torch.cuda.current_stream(dev).cuda_stream),
                "cmw_shard_kth")
    return out


def shard_merge(blocks, world: int, batch: int, k: int, k_out: int | None = None):
    """uint8[G * block_bytes] gathered shard blocks -> (scores f32[B,k_out], ids i64, scores64 f64, flags i32)."""
    torch = _torch()
    k_out = k if k_out is None else k_out
    status = None
    if isinstance(blocks, DevicePtr):
        dev, ptr, have, status = blocks.device, blocks.ptr, blocks.nbytes, blocks.status
    else:
        dev, ptr, have = blocks.device, blocks.data_ptr(), blocks.numel()
    out_s = torch.empty((batch, k_out), dtype=torch.float32, device=dev)
    out_i = torch.empty((batch, k_out), dtype=torch.int64, device=dev)
    out_s64 = torch.empty((batch, k_out), dtype=torch.float64, device=dev)
    flags = torch.zeros((batch,), dtype=torch.int32, device=dev)
    if batch:
        assert have >= world * int(N.lib().cmw_shard_block_bytes(batch, k))
        N.check(N.lib().cmw_shard_merge_ex(ptr, world, batch, k, k_out, out_s.data_ptr(), out_i.data_ptr(),
                                           out_s64.data_ptr(), flags.data_ptr(),
                                           status.data_ptr() if status is not None else None,
                                           torch.cuda.current_stream(dev).cuda_stream), "cmw_shard_merge")
    return out_s, out_i, out_s64, flags
