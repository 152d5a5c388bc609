"""B200Store -- drop-in for the reference's ``ChromaStore`` on the dense-retrieval hot path.

Mirrors ``rag_engine/storage/vector_store.py`` of the reference method for method (same names,
keyword arguments, return shapes and "fewer than k when the collection is small" behaviour):

* ``similarity_search_async(query_embedding, k=5)``   (vector_store.py:54-66)
* ``add_async(texts, metadatas, ids=None, embeddings=None)``  (:68-82)
* ``get_any_doc_meta_async(where)`` (:84-91), ``get_by_kb_id_async(kb_id)`` (:93-100)
* ``delete_where_async(where)`` (:102-105), ``get_collection()`` (:44-52)

so it can be handed to ``RAGRetriever(vector_store=...)`` (retrieval/retriever.py:34-46) or to
``top_k_search_async`` (retrieval/vector_search.py:8-10) unchanged.  The vectors live in HBM inside
``libcmwdense.so``; this class only keeps the host sidecar (documents, metadata, id maps, the
kbId -> dense group table) and coalesces concurrent ``similarity_search_async`` awaits -- the S
per-segment searches the reference gathers at retriever.py:179-182 -- into one batched launch.

Additions over the reference surface: ``search`` (batched, returns ids + scores), ``query``
(Chroma-shaped dict with ``ids`` / ``distances`` = 1 - cosine), ``search_multivector``.
"""
from __future__ import annotations

import asyncio
import logging
import threading
import uuid
from dataclasses import dataclass
from typing import Any

import numpy as np

from .engine import DenseStore
from .kbid import group_key


log = logging.getLogger("cmw_rag_b200.store")


@dataclass
class RetrievedDoc:
    """Same shape as rag_engine/storage/vector_store.py:13-16."""

    page_content: str
    metadata: dict[str, Any]


def _match(meta: dict[str, Any], where: dict[str, Any] | None) -> bool:
    """Chroma ``where`` subset: equality (what the reference uses: vector_store.py:87,96,105)
    plus $eq/$ne/$in/$nin and $and/$or."""
    if not where:
        return True
    for key, cond in where.items():
        if key == "$and":
            if not all(_match(meta, w) for w in cond):
                return False
        elif key == "$or":
            if not any(_match(meta, w) for w in cond):
                return False
        elif isinstance(cond, dict):
            val = meta.get(key)
            for op, ref in cond.items():
                if op == "$eq" and not (key in meta and val == ref):
                    return False
                if op == "$ne" and (key in meta and val == ref):
                    return False
                if op == "$in" and not (key in meta and val in ref):
                    return False
                if op == "$nin" and (key in meta and val in ref):
                    return False
        else:
            if key not in meta or meta[key] != cond:
                return False
    return True


class B200Store:
    # metadata keys with a hash index for `where` equality (every other filter scans the sidecar)
    INDEXED_KEYS = ("doc_stable_id", "kbId", "stable_id", "source_file")

    def __init__(
        self,
        collection_name: str = "default",
        host: str | None = None,  # accepted for signature compatibility; unused (no server)
        port: int | None = None,
        *,
        dim: int | None = None,
        capacity: int = 1 << 17,  # rows reserved in HBM up front; the collection doubles when it fills up
        device: int = 0,
        metric: str = "cosine",  # vector_store.py:48-51 pins {"hnsw:space": "cosine"}
        mode: str = "f32",
        keep_f32: bool = True,
        keep_bf16: bool = True,  # keep the 16-bit tiles (tensor-core filter); their format is `tiles16`
        id_offset: int = 0,
        tiles16: str = "f16",
        batch_window_us: float = 250.0,  # bounded gather window of the micro-batching front-end (batcher.py)
        max_batch: int = 64,
        max_queue: int = 4096,
        auto_compact: float | None = None,  # compact() by itself once this fraction of the rows is tombstoned
    ):
        self.collection_name = collection_name
        self.host = host
        self.port = port
        self.metric = metric
        self.mode = mode
        self._dim = dim
        self._capacity = int(capacity)
        self._device = int(device)
        self._keep = (keep_f32, keep_bf16)
        self._tiles16 = tiles16
        self._id_offset = int(id_offset)
        self._dense: DenseStore | None = None
        self._lock = threading.RLock()
        # host sidecar, indexed by LOCAL row number
        self._ids: list[str] = []
        self._docs: list[str | None] = []
        self._metas: list[dict[str, Any] | None] = []
        self._alive: list[bool] = []
        self._row_of: dict[str, int] = {}
        self._n_alive = 0
        # equality index on the metadata keys the reference filters by (vector_store.py:87,96,105 as driven
        # by core/indexer.py:397,432,505 and scripts/build_index.py:161-182): value -> rows in append order
        self._eq_index: dict[str, dict[Any, list[int]]] = {key: {} for key in self.INDEXED_KEYS}
        self._gid_of_key: dict[str, int] = {}
        self._key_of_gid: list[str] = []
        # micro-batching of concurrent callers: awaits of one event-loop tick are handed over as a group, and the
        # batcher (its own dispatcher thread, bounded window / batch / queue, histograms) merges groups, ticks,
        # threads and requests into launches
        self._pending: list[tuple[np.ndarray, int, asyncio.Future]] = []
        self._flush_scheduled = False
        self._batch_cfg = (int(max_batch), float(batch_window_us), int(max_queue))
        self._batcher_obj = None
        self._auto_compact = auto_compact
        self._uncertified = 0

    # -- plumbing -----------------------------------------------------------------------------
    def _ensure(self, dim: int) -> DenseStore:
        if self._dense is None:
            self._dim = int(dim)
            # the kernels want rows of 16-byte multiples; zero padding changes neither dots nor norms
            # (the reference's own store test uses 3-d vectors: tests/test_storage_vector_store.py:10-24)
            self._pdim = (self._dim + 7) // 8 * 8
            self._dense = DenseStore(self._pdim, self._capacity, device=self._device, f32=self._keep[0],
                                     bf16=self._keep[1], id_offset=self._id_offset, tiles16=self._tiles16)
        elif dim != self._dim:
            raise ValueError(f"embedding dimension {dim} does not match the collection's {self._dim}")
        return self._dense

    def _pad(self, x: np.ndarray) -> np.ndarray:
        if x.shape[1] == self._pdim:
            return x
        out = np.zeros((x.shape[0], self._pdim), np.float32)
        out[:, : x.shape[1]] = x
        return out

    @property
    def dense(self) -> DenseStore | None:
        return self._dense

    def count(self) -> int:
        return self._n_alive

    @property
    def batcher(self):
        """The micro-batching front-end of this collection (created on first use)."""
        if self._batcher_obj is None:
            with self._lock:
                if self._batcher_obj is None:
                    from .batcher import SearchBatcher

                    mb, win, mq = self._batch_cfg
                    self._batcher_obj = SearchBatcher(lambda q, kmax: self.search(q, kmax), max_batch=mb,
                                                      max_wait_us=win, max_queue=mq,
                                                      name=f"b200store_{self.collection_name}")
        return self._batcher_obj

    @property
    def stats(self) -> dict:
        b = self._batcher_obj
        c = b.counters if b is not None else {"requests": 0, "launches": 0}
        return {"searches": c["requests"], "launch_batches": c["launches"],
                "max_batch": int(b.h_batch.max) if b is not None else 0, "uncertified": self._uncertified}

    def metrics(self) -> dict:
        """Queue depth / batch size / latency histograms of the seam + collection gauges."""
        m = self.batcher.metrics()
        m["collection"] = {"name": self.collection_name, "rows": len(self._ids), "live_rows": self._n_alive,
                           "uncertified_queries": self._uncertified}
        return m

    def close(self) -> None:
        if self._batcher_obj is not None:
            self._batcher_obj.close()
            self._batcher_obj = None
        if self._dense is not None:
            self._dense.close()
            self._dense = None

    def _index_row(self, row: int, meta: dict[str, Any] | None) -> None:
        if not meta:
            return
        for key in self.INDEXED_KEYS:
            if key in meta:
                try:
                    self._eq_index[key].setdefault(meta[key], []).append(row)
                except TypeError:  # unhashable value: only reachable through the full scan
                    pass

    def _index_candidates(self, where) -> list[int] | None:
        """Rows that can match `where` according to the equality index (None = no indexed condition:
        scan).  Only top-level `key: value` / `key: {"$eq": value}` conditions are used; the caller still
        applies the whole filter."""
        if not isinstance(where, dict):
            return None
        best = None
        for key, cond in where.items():
            if key not in self._eq_index:
                continue
            if isinstance(cond, dict):
                if set(cond) != {"$eq"}:
                    continue
                cond = cond["$eq"]
            try:
                rows = self._eq_index[key].get(cond, [])
            except TypeError:
                continue
            if best is None or len(rows) < len(best):
                best = rows
        return best

    def gid_for(self, raw_kb_id) -> int:
        """Dense group number of a kbId (retriever.py:236-239 grouping key), -1 for a falsy kbId."""
        key = group_key(raw_kb_id)
        if key is None:
            return -1
        gid = self._gid_of_key.get(key)
        if gid is None:
            gid = len(self._key_of_gid)
            self._gid_of_key[key] = gid
            self._key_of_gid.append(key)
        return gid

    def key_of_gid(self, gid: int) -> str:
        return self._key_of_gid[gid]

    # -- ingest / mutation (sync cores + the reference's async names) ------------------------------
    def add(self, texts, metadatas, ids=None, embeddings=None) -> None:
        if embeddings is None:
            raise ValueError("B200Store.add needs embeddings (the reference's indexer always passes "
                             "them: rag_engine/core/indexer.py:495-507); there is no embedding function")
        n = len(texts)
        emb = np.ascontiguousarray(embeddings, dtype=np.float32)
        if emb.ndim != 2 or emb.shape[0] != n:
            raise ValueError(f"embeddings must be [{n}, dim], got {emb.shape}")
        metadatas = list(metadatas) if metadatas is not None else [None] * n
        if len(metadatas) != n:
            raise ValueError("texts and metadatas differ in length")
        ids = [str(i) for i in ids] if ids is not None else [uuid.uuid4().hex for _ in range(n)]
        if len(ids) != n or len(set(ids)) != n:
            raise ValueError("ids must be unique and match texts in length")
        with self._lock:
            keep = [i for i in range(n) if ids[i] not in self._row_of]  # Chroma add() ignores known ids
            if not keep:
                return
            dense = self._ensure(emb.shape[1])
            gids = np.array([self.gid_for((metadatas[i] or {}).get("kbId", "")) for i in keep], np.int32)
            base = len(self._ids)
            if base + len(keep) > self._capacity:
                dense = self._grow(base + len(keep))
            dense.append(self._pad(emb[keep]), gids)
            for j, i in enumerate(keep):
                self._row_of[ids[i]] = base + j
                self._ids.append(ids[i])
                self._docs.append(texts[i])
                self._metas.append(dict(metadatas[i]) if metadatas[i] is not None else None)
                self._alive.append(True)
                self._index_row(base + j, metadatas[i])
            self._n_alive += len(keep)

    def _grow(self, needed: int) -> DenseStore:
        """A Chroma collection grows without bound; the HBM store is sized up front.  When an add would
        overflow it, a store twice as large is created and the rows are re-ingested device to device
        (cmw_store_copy_rows: no host round trip), tombstones replayed.  Needs the fp32 tiles."""
        if not self._keep[0]:
            raise RuntimeError(f"collection is full ({self._capacity} rows) and keeps no fp32 tiles to grow from")
        new_cap = max(needed, 2 * self._capacity)
        old = self._dense
        new = DenseStore(self._pdim, new_cap, device=self._device, f32=self._keep[0], bf16=self._keep[1],
                         id_offset=self._id_offset, tiles16=self._tiles16)
        n = len(self._ids)
        new.copy_rows_from(old, row0=0, n=n)
        dead = [r for r, alive in enumerate(self._alive) if not alive]
        if dead:
            new.tombstone(dead)
        old.close()
        self._dense = new
        self._capacity = new_cap
        return new

    def _rows_where(self, where, limit: int | None = None) -> list[int]:
        cand = self._index_candidates(where)
        rows = cand if cand is not None else range(len(self._ids))
        out = []
        for r in rows:
            if self._alive[r] and _match(self._metas[r] or {}, where):
                out.append(r)
                if limit is not None and len(out) >= limit:
                    break
        return out

    def compact(self) -> int:
        """Drop tombstoned rows physically (they cost scan bandwidth until then): the live rows are re-ingested,
        in order and device to device (cmw_store_copy_rows with a gather index), into a fresh HBM store and the
        sidecar is renumbered.
        Row numbers returned by :meth:`search` change; string ids, documents and metadata do not.  Returns
        the number of rows reclaimed.  Needs the fp32 tiles."""
        with self._lock:
            n = len(self._ids)
            dead = n - self._n_alive
            if dead == 0 or self._dense is None:
                return 0
            if not self._keep[0]:
                raise RuntimeError("compact() re-ingests from the fp32 tiles, which this collection does not keep")
            old = self._dense
            new = DenseStore(self._pdim, self._capacity, device=self._device, f32=self._keep[0], bf16=self._keep[1],
                             id_offset=self._id_offset, tiles16=self._tiles16)
            alive = np.asarray(self._alive, bool)
            new.copy_rows_from(old, rows=np.flatnonzero(alive))
            old.close()
            self._dense = new
            live_rows = np.flatnonzero(alive).tolist()
            self._ids = [self._ids[r] for r in live_rows]
            self._docs = [self._docs[r] for r in live_rows]
            self._metas = [self._metas[r] for r in live_rows]
            self._alive = [True] * len(live_rows)
            self._row_of = {sid: r for r, sid in enumerate(self._ids)}
            self._eq_index = {key: {} for key in self.INDEXED_KEYS}
            for r, meta in enumerate(self._metas):
                self._index_row(r, meta)
            return dead

    def delete(self, where=None, ids=None) -> int:
        with self._lock:
            rows = set()
            if ids is not None:
                rows.update(self._row_of[i] for i in ids if i in self._row_of)
                if where:
                    rows = {r for r in rows if _match(self._metas[r] or {}, where)}
            elif where:
                rows.update(self._rows_where(where))
            rows = sorted(r for r in rows if self._alive[r])
            if rows and self._dense is not None:
                self._dense.tombstone(rows)
            for r in rows:
                self._alive[r] = False
                self._row_of.pop(self._ids[r], None)
                self._docs[r] = None
                self._metas[r] = None  # its index entries stay behind and are skipped by the alive check
            self._n_alive -= len(rows)
            n_all = len(self._ids)
            if (self._auto_compact is not None and n_all >= 4096 and self._keep[0]
                    and n_all - self._n_alive > self._auto_compact * n_all):
                self.compact()  # tombstones cost scan bandwidth and thin out the first slabs: reclaim them
            return len(rows)

    def get(self, where=None, ids=None, include=("metadatas", "documents"), limit=None) -> dict:
        with self._lock:
            if ids is not None:
                rows = [self._row_of[i] for i in ids if i in self._row_of]
                rows = [r for r in rows if _match(self._metas[r] or {}, where)]
                if limit is not None:
                    rows = rows[:limit]
            else:
                rows = self._rows_where(where, limit)
            out: dict[str, Any] = {"ids": [self._ids[r] for r in rows]}
            out["metadatas"] = [self._metas[r] for r in rows] if "metadatas" in include else None
            out["documents"] = [self._docs[r] for r in rows] if "documents" in include else None
            return out

    async def add_async(self, texts, metadatas, ids=None, embeddings=None) -> None:
        await asyncio.to_thread(self.add, texts, metadatas, ids, embeddings)

    async def get_any_doc_meta_async(self, where: dict[str, Any]) -> dict[str, Any] | None:
        metas = self.get(where=where, include=["metadatas"], limit=1).get("metadatas") or []
        return metas[0] if metas else None

    async def get_by_kb_id_async(self, kb_id: str) -> dict[str, Any] | None:
        metas = self.get(where={"kbId": kb_id}, include=["metadatas"], limit=1).get("metadatas") or []
        return metas[0] if metas else None

    async def delete_where_async(self, where: dict[str, Any]) -> None:
        await asyncio.to_thread(self.delete, where)

    async def get_collection(self):
        return _Collection(self)

    # -- search ---------------------------------------------------------------------------------
    def search(self, queries, k: int, mode: str | None = None, algo: str | None = None):
        """Batched search from HOST vectors: (scores f32[B,k'], ids i64[B,k'], flags) numpy, where
        slots beyond the number of live rows hold id -1 / score -inf."""
        q = np.ascontiguousarray(np.atleast_2d(np.asarray(queries, dtype=np.float32)))
        with self._lock:
            if self._dense is None or q.shape[0] == 0:
                b = q.shape[0]
                return (np.full((b, k), -np.inf, np.float32), np.full((b, k), -1, np.int64),
                        np.zeros((b,), np.int32))
            if q.shape[1] != self._dim:
                raise ValueError(f"query dimension {q.shape[1]} does not match the collection's {self._dim}")
            out = self._dense.search_host(self._pad(q), k, metric=self.metric, mode=mode or self.mode,
                                          algo=algo)
        self._note_flags(out[2])
        return out

    def _note_flags(self, flags) -> None:
        """A query still flagged after the library's repair chain (in practice: more exact duplicates of a
        top-k chunk than the candidate set holds) is answered with the best candidates found, but it is not
        PROVEN identical to the exact answer: count it and say so -- never pass it on silently."""
        bad = int(np.count_nonzero(flags))
        if bad:
            self._uncertified += bad
            log.warning("B200Store[%s]: %d of %d queries came back CMW_FLAG_UNCERTIFIED after the repair chain "
                        "(candidate pool overflow or unbreakable ties); results are the best rescored candidates",
                        self.collection_name, bad, len(flags))

    def _docs_for(self, ids_row: np.ndarray) -> list[RetrievedDoc]:
        out = []
        off = self._id_offset
        with self._lock:
            docs, metas, alive, n = self._docs, self._metas, self._alive, len(self._ids)
            for gid in ids_row.tolist():
                r = gid - off
                if gid >= 0 and 0 <= r < n and alive[r]:
                    m = metas[r]
                    out.append(RetrievedDoc(docs[r], m.copy() if m else {}))
        return out

    def query(self, query_embeddings, n_results: int = 10, include=("documents", "metadatas", "distances")):
        """Chroma-shaped result dict: every value is a list (one per query) of lists."""
        scores, ids, _ = self.search(query_embeddings, n_results)
        res: dict[str, Any] = {"ids": [], "documents": None, "metadatas": None, "distances": None}
        docs, metas, dists = [], [], []
        for b in range(ids.shape[0]):
            # ids and distances are filtered by the SAME list of kept slots, so they always line up
            kept = [j for j, g in enumerate(ids[b].tolist())
                    if g >= 0 and 0 <= g - self._id_offset < len(self._ids) and self._alive[g - self._id_offset]]
            rows = [int(ids[b, j]) - self._id_offset for j in kept]
            res["ids"].append([self._ids[r] for r in rows])
            docs.append([self._docs[r] for r in rows])
            metas.append([dict(self._metas[r] or {}) for r in rows])
            dists.append([float(1.0 - scores[b, j]) for j in kept])
        if "documents" in include:
            res["documents"] = docs
        if "metadatas" in include:
            res["metadatas"] = metas
        if "distances" in include:
            res["distances"] = dists
        return res

    def similarity_search(self, query_embedding, k: int = 5) -> list[RetrievedDoc]:
        """Blocking form; concurrent callers (threads) share launches through the batcher."""
        _, ids, _ = self.batcher.search_one(np.asarray(query_embedding, dtype=np.float32), int(k))
        return self._docs_for(ids)

    async def similarity_search_async(self, query_embedding: list[float], k: int = 5) -> list[RetrievedDoc]:
        """Same contract as ChromaStore.similarity_search_async (vector_store.py:54-66).  Awaits issued in the
        same event-loop tick (asyncio.gather over segments, retriever.py:179-182) are handed to the batcher as one
        group; the batcher merges groups from different ticks, threads and requests inside its bounded window.
        The event loop is never blocked: the launch runs on the batcher's thread."""
        loop = asyncio.get_running_loop()
        fut: asyncio.Future = loop.create_future()
        self._pending.append((np.asarray(query_embedding, dtype=np.float32), int(k), fut))
        if not self._flush_scheduled:
            self._flush_scheduled = True
            loop.call_soon(lambda: asyncio.ensure_future(self._flush()))
        return await fut

    async def _flush(self) -> None:
        from .batcher import QueueFull

        group, self._pending = self._pending, []
        self._flush_scheduled = False
        if not group:
            return
        try:
            futs = []
            step = max(1, self.batcher.max_queue // 2)  # a tick's group larger than the queue goes in pieces
            for lo in range(0, len(group), step):
                part = group[lo:lo + step]
                while True:
                    try:
                        futs += self.batcher.submit_many([v for v, _, _ in part], [k for _, k, _ in part])
                        break
                    except QueueFull:  # back-pressure without blocking the loop
                        await asyncio.sleep(0.001)
            results = await asyncio.gather(*[asyncio.wrap_future(f) for f in futs], return_exceptions=True)
            for (_, _, fut), res in zip(group, results):
                if fut.done():
                    continue
                if isinstance(res, BaseException):  # propagate like a chromadb error would (retrieve_context.py:435-449)
                    fut.set_exception(res)
                else:
                    fut.set_result(self._docs_for(res[1]))
        except Exception as exc:
            for _, _, fut in group:
                if not fut.done():
                    fut.set_exception(exc)

    # -- persistence (the analogue of the Chroma server's --path directory) ----------
    FORMAT_VERSION = 1

    def save(self, path: str, chunk_rows: int = 65536) -> None:
        """Write the collection to a directory: meta.json, rows.f32 (raw little-endian [rows, dim],
        as appended), kb_gid.i32, alive.u8 and sidecar.jsonl (id, document, metadata per row)."""
        import json
        import os

        os.makedirs(path, exist_ok=True)
        with self._lock:
            n = len(self._ids)
            meta = {"format": self.FORMAT_VERSION, "collection_name": self.collection_name, "metric": self.metric,
                    "mode": self.mode, "dim": self._dim, "padded_dim": getattr(self, "_pdim", None), "rows": n,
                    "id_offset": self._id_offset, "group_keys": self._key_of_gid}
            with open(os.path.join(path, "rows.f32"), "wb") as fr, open(os.path.join(path, "kb_gid.i32"), "wb") as fg:
                for lo in range(0, n, chunk_rows):
                    m = min(chunk_rows, n - lo)
                    rows, gid, _ = self._dense.read_rows(lo, m)
                    fr.write(np.ascontiguousarray(rows[:, : self._dim]).tobytes())
                    fg.write(gid.tobytes())
            np.asarray(self._alive, np.uint8).tofile(os.path.join(path, "alive.u8"))
            with open(os.path.join(path, "sidecar.jsonl"), "w", encoding="utf-8") as f:
                for i in range(n):
                    f.write(json.dumps({"id": self._ids[i], "document": self._docs[i], "metadata": self._metas[i]},
                                       ensure_ascii=False) + "\n")
            with open(os.path.join(path, "meta.json"), "w") as f:  # written last: marks a complete save
                json.dump(meta, f)

    @classmethod
    def load(cls, path: str, capacity: int | None = None, device: int = 0, chunk_rows: int = 65536, **kw) -> "B200Store":
        """Rebuild a collection saved by save(): rows stream from the memory-mapped file through the
        pinned staging buffer into HBM (K0 recomputes the bf16 tiles and norms), tombstones are replayed."""
        import json
        import os

        with open(os.path.join(path, "meta.json")) as f:
            meta = json.load(f)
        if meta.get("format") != cls.FORMAT_VERSION:
            raise ValueError(f"unsupported store format {meta.get('format')!r}")
        n, dim = int(meta["rows"]), meta["dim"]
        store = cls(meta["collection_name"], dim=dim, capacity=max(int(capacity or 0), n, 1), device=device,
                    metric=meta["metric"], mode=meta["mode"], id_offset=int(meta.get("id_offset", 0)), **kw)
        store._key_of_gid = list(meta["group_keys"])
        store._gid_of_key = {k: i for i, k in enumerate(store._key_of_gid)}
        if n == 0:
            return store
        rows = np.memmap(os.path.join(path, "rows.f32"), dtype=np.float32, mode="r", shape=(n, dim))
        gid = np.fromfile(os.path.join(path, "kb_gid.i32"), dtype=np.int32)
        alive = np.fromfile(os.path.join(path, "alive.u8"), dtype=np.uint8).astype(bool)
        dense = store._ensure(dim)
        for lo in range(0, n, chunk_rows):
            hi = min(n, lo + chunk_rows)
            dense.append(store._pad(np.ascontiguousarray(rows[lo:hi])), gid[lo:hi])
        with open(os.path.join(path, "sidecar.jsonl"), encoding="utf-8") as f:
            for i, line in enumerate(f):
                rec = json.loads(line)
                store._ids.append(rec["id"])
                store._docs.append(rec["document"])
                store._metas.append(rec["metadata"])
                store._alive.append(bool(alive[i]))
                if alive[i]:
                    store._row_of[rec["id"]] = i
                    store._index_row(i, rec["metadata"])
                    store._n_alive += 1
        dead = np.flatnonzero(~alive)
        if dead.size:
            dense.tombstone(dead)
        return store

    def search_multivector(self, segment_embeddings, k: int, prl: int = 0, limit: int = 0):
        """[Q, S, dim] host array -> (MultiVectorResult on CPU, ids [Q,S,k], scores [Q,S,k])."""
        import torch

        seg = np.ascontiguousarray(segment_embeddings, dtype=np.float32)
        with self._lock:
            if self._dense is None:
                raise RuntimeError("empty collection")
            dev = torch.device(f"cuda:{self._device}")
            t = torch.from_numpy(seg).to(dev)
            res, scores, ids, flags = self._dense.search_multivector(t, k, prl=prl, limit=limit,
                                                                     metric=self.metric, mode=self.mode)
            torch.cuda.synchronize(dev)
            flags = flags.cpu().numpy().ravel()
            if flags.any():
                # the device path has no repair chain: re-run the flagged segments through the host API (which
                # has), then redo the reduction on the repaired per-segment lists
                qn, sn, _ = seg.shape
                bad = np.flatnonzero(flags)
                sc_r, ids_r, fl_r = self._dense.search_host(self._pad(seg.reshape(qn * sn, -1)[bad]), k,
                                                            metric=self.metric, mode=self.mode)
                ids2, sc2 = ids.reshape(qn * sn, k).clone(), scores.reshape(qn * sn, k).clone()
                ids2[torch.from_numpy(bad).to(dev)] = torch.from_numpy(ids_r).to(dev)
                sc2[torch.from_numpy(bad).to(dev)] = torch.from_numpy(sc_r).to(dev)
                ids, scores = ids2.view(qn, sn, k), sc2.view(qn, sn, k)
                res = self._dense.multivector(ids, scores, prl=prl, limit=limit)
                torch.cuda.synchronize(dev)
                self._note_flags(fl_r)
            return res.cpu(), ids.cpu().numpy(), scores.cpu().numpy()


class _Collection:
    """The slice of chromadb's AsyncCollection API the reference touches."""

    def __init__(self, store: B200Store):
        self._s = store
        self.name = store.collection_name
        self.metadata = {"hnsw:space": store.metric}

    async def query(self, query_embeddings, n_results: int = 10, include=("documents", "metadatas", "distances"),
                    **_ignored):
        return await asyncio.to_thread(self._s.query, query_embeddings, n_results, include)

    async def add(self, ids=None, documents=None, metadatas=None, embeddings=None):
        await self._s.add_async(documents, metadatas, ids=ids, embeddings=embeddings)

    async def get(self, where=None, ids=None, include=("metadatas", "documents"), limit=None, **_ignored):
        return self._s.get(where=where, ids=ids, include=include, limit=limit)

    async def delete(self, where=None, ids=None):
        await asyncio.to_thread(self._s.delete, where, ids)

    async def count(self) -> int:
        return self._s.count()
