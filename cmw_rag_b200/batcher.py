"""Micro-batching front-end for concurrent ``similarity_search_async`` callers.

The reference issues one store call per query vector: S concurrent awaits per request
(rag_engine/retrieval/retriever.py:179-182) times the concurrent requests the UI allows
(rag_engine/tools/retrieve_context.py:397-409, config/settings.py:166).  On the GPU a batch of 16 costs what a
batch of 1 costs (the scan is HBM-bound), so requests that arrive a few hundred microseconds apart should share
one launch.  ``SearchBatcher`` does that behind any ``search(queries[B, D], k) -> (scores, ids, flags)`` callable:

* a request waits at most ``max_wait_us`` after it was queued (bounded gather window) and a launch takes at most
  ``max_batch`` requests;
* while a launch is running, new requests simply queue up: under load the batch size follows the arrival rate by
  itself and the window never adds latency (the oldest request is already older than the window when the
  dispatcher comes back);
* back-pressure: at most ``max_queue`` requests may wait; ``submit`` then blocks (``block=True``) or raises
  ``QueueFull``;
* metrics: histograms of queue depth at dispatch, batch size, time in queue, launch time and end-to-end seam
  latency, as a dict (``metrics()``) and in Prometheus text exposition format (``prometheus()``).  The reference
  has no instrumentation around retrieval at all.
"""
from __future__ import annotations

import bisect
import concurrent.futures
import threading
import time
from collections import deque

import numpy as np


class QueueFull(RuntimeError):
    """submit(block=False) on a batcher whose queue holds ``max_queue`` requests."""


class Histogram:
    """Fixed-bucket histogram (cumulative on export, like Prometheus)."""

    def __init__(self, bounds):
        self.bounds = list(bounds)
        self.counts = [0] * (len(self.bounds) + 1)
        self.total = 0
        self.sum = 0.0
        self.max = 0.0

    def observe(self, v: float) -> None:
        self.counts[bisect.bisect_left(self.bounds, v)] += 1
        self.total += 1
        self.sum += v
        if v > self.max:
            self.max = v

    def quantile(self, q: float) -> float:
        """Upper bucket bound that covers quantile q (the max for the overflow bucket)."""
        if not self.total:
            return 0.0
        need, run = q * self.total, 0
        for i, c in enumerate(self.counts):
            run += c
            if run >= need:
                return self.bounds[i] if i < len(self.bounds) else self.max
        return self.max

    def as_dict(self) -> dict:
        return {"count": self.total, "sum": self.sum, "mean": self.sum / self.total if self.total else 0.0,
                "max": self.max, "p50": self.quantile(0.5), "p99": self.quantile(0.99),
                "buckets": {str(b): c for b, c in zip(self.bounds + ["+Inf"], self.counts)}}

    def prometheus(self, name: str, help_: str) -> str:
        out = [f"# HELP {name} {help_}", f"# TYPE {name} histogram"]
        run = 0
        for b, c in zip(self.bounds, self.counts):
            run += c
            out.append(f'{name}_bucket{{le="{b}"}} {run}')
        out.append(f'{name}_bucket{{le="+Inf"}} {self.total}')
        out.append(f"{name}_sum {self.sum}")
        out.append(f"{name}_count {self.total}")
        return "\n".join(out)


_MS = [0.05, 0.1, 0.2, 0.35, 0.5, 0.75, 1, 1.5, 2, 3, 5, 10, 20, 50, 100, 1000]
_SIZES = [1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 4096]


class SearchBatcher:
    def __init__(self, search, max_batch: int = 64, max_wait_us: float = 200.0, max_queue: int = 4096,
                 name: str = "b200store"):
        """``search``: callable (queries f32[B, D], kmax) -> (scores [B, kmax], ids [B, kmax], flags [B]), run on
        the dispatcher thread, one call at a time."""
        self._search = search
        self.max_batch = int(max_batch)
        self.max_wait = float(max_wait_us) * 1e-6
        self.max_queue = int(max_queue)
        self.name = name
        self._q: deque = deque()
        self._cv = threading.Condition()
        self._closed = False
        self.h_batch = Histogram(_SIZES)
        self.h_depth = Histogram(_SIZES)
        self.h_wait_ms = Histogram(_MS)
        self.h_launch_ms = Histogram(_MS)
        self.h_seam_ms = Histogram(_MS)
        self.counters = {"requests": 0, "launches": 0, "rejected": 0, "errors": 0, "flagged": 0}
        self._thread = threading.Thread(target=self._run, name=f"{name}-batcher", daemon=True)
        self._thread.start()

    # -- producer side ------------------------------------------------------------------------------
    def submit(self, vector, k: int, block: bool = True, timeout: float | None = None) -> concurrent.futures.Future:
        """Queue one query; the future resolves to (scores f32[k], ids i64[k], flag)."""
        fut: concurrent.futures.Future = concurrent.futures.Future()
        v = np.asarray(vector, dtype=np.float32).reshape(-1)
        with self._cv:
            if self._closed:
                raise RuntimeError("SearchBatcher is closed")
            if len(self._q) >= self.max_queue:
                if not block:
                    self.counters["rejected"] += 1
                    raise QueueFull(f"{self.name}: {len(self._q)} requests are waiting (max_queue {self.max_queue})")
                if not self._cv.wait_for(lambda: len(self._q) < self.max_queue or self._closed, timeout):
                    self.counters["rejected"] += 1
                    raise QueueFull(f"{self.name}: queue still full after {timeout} s")
                if self._closed:
                    raise RuntimeError("SearchBatcher is closed")
            self._q.append((v, int(k), fut, time.perf_counter()))
            self.counters["requests"] += 1
            self._cv.notify_all()
        return fut

    def submit_many(self, vectors, ks) -> list:
        """Queue several queries under one lock acquisition (the S segment searches one request gathers): they sit
        next to each other in the queue and share a launch unless ``max_batch`` splits them.  Never blocks: raises
        ``QueueFull`` if they do not all fit."""
        now = time.perf_counter()
        futs = [concurrent.futures.Future() for _ in ks]
        with self._cv:
            if self._closed:
                raise RuntimeError("SearchBatcher is closed")
            if len(self._q) + len(futs) > self.max_queue:
                self.counters["rejected"] += len(futs)
                raise QueueFull(f"{self.name}: {len(self._q)} requests are waiting (max_queue {self.max_queue})")
            for v, k, fut in zip(vectors, ks, futs):
                self._q.append((np.asarray(v, dtype=np.float32).reshape(-1), int(k), fut, now))
            self.counters["requests"] += len(futs)
            self._cv.notify_all()
        return futs

    def search_one(self, vector, k: int):
        return self.submit(vector, k).result()

    # -- dispatcher ----------------------------------------------------------------------------------
    def _take(self):
        with self._cv:
            while not self._q and not self._closed:
                self._cv.wait()
            if not self._q:
                return None
            # the gather window is measured from the OLDEST waiting request: bounded added latency
            deadline = self._q[0][3] + self.max_wait
            while len(self._q) < self.max_batch and not self._closed:
                left = deadline - time.perf_counter()
                if left <= 0:
                    break
                self._cv.wait(left)
            depth = len(self._q)
            batch = [self._q.popleft() for _ in range(min(depth, self.max_batch))]
            self.h_depth.observe(depth)
            self._cv.notify_all()  # room for blocked producers
            return batch

    def _run(self):
        while True:
            batch = self._take()
            if batch is None:
                return
            t0 = time.perf_counter()
            for _, _, _, t_in in batch:
                self.h_wait_ms.observe((t0 - t_in) * 1e3)
            self.h_batch.observe(len(batch))
            try:
                dims = {v.shape[0] for v, _, _, _ in batch}
                if len(dims) != 1:
                    raise ValueError(f"queries of different dimensions in one batch: {sorted(dims)}")
                kmax = max(k for _, k, _, _ in batch)
                scores, ids, flags = self._search(np.stack([v for v, _, _, _ in batch]), kmax)
                t1 = time.perf_counter()
                self.h_launch_ms.observe((t1 - t0) * 1e3)
                self.counters["launches"] += 1
                self.counters["flagged"] += int(np.count_nonzero(flags))
                for i, (_, k, fut, t_in) in enumerate(batch):
                    self.h_seam_ms.observe((t1 - t_in) * 1e3)
                    if not fut.cancelled():
                        fut.set_result((scores[i, :k], ids[i, :k], int(flags[i])))
            except Exception as exc:  # propagate to every waiter, like a chromadb error would reach each await
                self.counters["errors"] += 1
                for _, _, fut, _ in batch:
                    if not fut.done():
                        fut.set_exception(exc)

    # -- lifecycle / metrics ---------------------------------------------------------------------------
    def close(self) -> None:
        with self._cv:
            self._closed = True
            self._cv.notify_all()
        self._thread.join(timeout=30)
        with self._cv:
            while self._q:
                _, _, fut, _ = self._q.popleft()
                if not fut.done():
                    fut.set_exception(RuntimeError("SearchBatcher closed"))

    def queue_depth(self) -> int:
        with self._cv:
            return len(self._q)

    def metrics(self) -> dict:
        return {"counters": dict(self.counters), "queue_depth_now": self.queue_depth(),
                "batch_size": self.h_batch.as_dict(), "queue_depth_at_dispatch": self.h_depth.as_dict(),
                "queue_wait_ms": self.h_wait_ms.as_dict(), "launch_ms": self.h_launch_ms.as_dict(),
                "seam_latency_ms": self.h_seam_ms.as_dict(),
                "config": {"max_batch": self.max_batch, "max_wait_us": self.max_wait * 1e6, "max_queue": self.max_queue}}

    def prometheus(self) -> str:
        p = f"cmw_{self.name}"
        parts = [self.h_batch.prometheus(f"{p}_batch_size", "queries per kernel launch"),
                 self.h_depth.prometheus(f"{p}_queue_depth", "requests waiting when a launch was dispatched"),
                 self.h_wait_ms.prometheus(f"{p}_queue_wait_ms", "milliseconds a request waited for its launch"),
                 self.h_launch_ms.prometheus(f"{p}_launch_ms", "milliseconds per batched search call"),
                 self.h_seam_ms.prometheus(f"{p}_seam_latency_ms", "milliseconds from submit to result")]
        for name, val in self.counters.items():
            parts.append(f"# TYPE {p}_{name}_total counter\n{p}_{name}_total {val}")
        parts.append(f"# TYPE {p}_queue_depth_now gauge\n{p}_queue_depth_now {self.queue_depth()}")
        return "\n".join(parts) + "\n"
