#!/usr/bin/env python
"""bench.py -- queries/sec of the dense-retrieval hot path at 1M x 1536, top-100 (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the CPU arm (exact oracle port /
                                                             # hnswlib-equivalent HNSW), rank 0 only

A step = one batch of B query vectors through cmw_search (exact mode: tensor-core / scan filter,
fp64 rescoring, certificate) over the HBM-resident corpus.  `value` is device-resident throughput
(CUDA events on the launching stream); `e2e` goes through the host-buffer C-ABI call
(cmw_search_host: pinned staging, H2D, kernels, D2H, sync).  N > 1: one process per GPU
(torchrun), the corpus replicated and the query batch sharded (no data-path collective), or
`--shard rows` for the row-sharded corpus with an NCCL all-gather + merge kernel.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

METRIC = "queries/sec at 1Mx1536 top-100"
UNIT = "queries/s"
SEED = 20261018


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=1_000_000, help="corpus rows per replica (or per shard)")
    ap.add_argument("--dim", type=int, default=1536)
    ap.add_argument("--batch", type=int, default=4096, help="query vectors per step per GPU")
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--mode", default="f32", choices=["f32", "bf16"])
    ap.add_argument("--algo", default="auto")
    ap.add_argument("--shard", default="queries", choices=["queries", "rows"])
    ap.add_argument("--exchange", default="nccl", choices=["nccl", "peer"],
                    help="--shard rows: NCCL all-gather + merge kernel, or the fused NVLink peer-memory exchange")
    ap.add_argument("--no-f32", action="store_true", help="bf16 tiles only (rows-sharded 200M config)")
    ap.add_argument("--cpu-queries", type=int, default=16, help="queries in the bounded CPU sample")
    ap.add_argument("--hnsw-rows", type=int, default=0,
                    help="rows in the CPU HNSW index (bounded sample); 0 = 50k for the cpu_baseline leg of the "
                         "GPU arm (~10 s build), and for --impl reference as many rows as --hnsw-build-budget allows")
    ap.add_argument("--hnsw-build-budget", type=float, default=150.0,
                    help="--impl reference: seconds of index construction before the index is frozen")
    ap.add_argument("--hnsw-queries", type=int, default=512)
    ap.add_argument("--in-flight", type=int, default=2, help="requests outstanding in the pipelined end-to-end leg")
    ap.add_argument("--leg-gap", type=float, default=1.0, help="idle seconds before each timed leg")
    ap.add_argument("--skip-cpu-exact", action="store_true")
    ap.add_argument("--skip-b1", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md: sample nvidia-smi DURING the timed region)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows: list[list[str]] = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) >= 7:
                self.rows.append(parts)

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()  # the exact PID we started
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except ValueError:
                continue
            for name, val in zip(names, r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks() -> dict:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


# ------------------------------------------------------------------------------------------------
# CPU arm
# ------------------------------------------------------------------------------------------------
def host_corpus(rows: int, dim: int) -> np.ndarray:
    import synth

    return synth.make_corpus(rows, dim, seed=SEED, ties=False)


def cpu_exact_sample(corpus: np.ndarray, queries: np.ndarray, k: int, steps: int = 1):
    """Times the oracle's C port (exact fp64 top-k, OpenMP over all host threads)."""
    from oracle.cport import exact_topk_c, num_threads

    exact_topk_c(corpus[:4096], queries[:1], k)  # page in / thread pool warm-up
    t0 = time.perf_counter()
    for _ in range(steps):
        exact_topk_c(corpus, queries, k)
    dt = time.perf_counter() - t0
    return queries.shape[0] * steps / dt, dt, num_threads()


def cpu_exact_sgemm_sample(corpus: np.ndarray, queries: np.ndarray, k: int):
    """Exact fp32 top-k the way a CPU library does it: one sgemm ``Q @ C.T`` on all host
    threads (torch CPU = MKL / oneDNN) and a partial sort per query.  The corpus rows are unit vectors here,
    so the dot product is the cosine.  Returns (qps, seconds, threads, ids)."""
    import torch

    c = torch.from_numpy(corpus)
    q = torch.from_numpy(np.ascontiguousarray(queries))
    torch.topk(q[:2] @ c[:4096].T, min(k, 4096), dim=1)  # thread pool / kernel selection warm-up
    t0 = time.perf_counter()
    scores = q @ c.T
    _, ids = torch.topk(scores, k, dim=1, sorted=True)
    dt = time.perf_counter() - t0
    return queries.shape[0] / dt, dt, torch.get_num_threads(), ids.numpy()


def cpu_hnsw_sample(corpus: np.ndarray, queries: np.ndarray, k: int, index_rows: int, steps: int = 1):
    """The "Chroma HNSW" baseline (BASELINE.json north_star): an hnswlib-equivalent index with
    Chroma's defaults (M=16, ef_construction=100, ef_search=100) over the first `index_rows` rows,
    all host threads, one query per thread.  Returns (qps, recall@k vs exact on the same rows,
    build seconds, threads)."""
    from oracle.cport import exact_topk_c
    from oracle.hnsw import HnswIndex, num_threads

    sub = corpus[:index_rows]
    ix = HnswIndex(sub.shape[1], sub.shape[0])
    t0 = time.perf_counter()
    ix.add(sub)
    build_s = time.perf_counter() - t0
    ix.search(queries[:8], k)
    t0 = time.perf_counter()
    for _ in range(steps):
        ids, _ = ix.search(queries, k)
    dt = time.perf_counter() - t0
    nref = min(64, queries.shape[0])
    ref, _, _ = exact_topk_c(sub, queries[:nref], k)
    recall = float(np.mean([len(set(ids[b]) & set(ref[b])) / k for b in range(nref)]))
    ix.close()
    return queries.shape[0] * steps / dt, recall, build_s, num_threads()


HNSW_LABEL = ("hnswlib-equivalent HNSW re-implementation (oracle/hnsw/hnsw_baseline.cpp), chromadb 1.3.0 "
              "defaults assumed (M=16, ef_construction=100, ef_search=100), not verifiable offline")


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on the host cores.  The
    reference delegates the arithmetic to chromadb/hnswlib (not installable here); what is timed is the
    hnswlib-equivalent HNSW index of oracle/hnsw (Chroma's defaults, all host threads, one query per
    thread) answering a bounded sample of queries per step.  The index covers as much of the corpus as
    can be inserted within --hnsw-build-budget seconds (a full 1M x 1536 build takes ~250 s on 16
    cores; HNSW query cost grows ~log N, so a partial index flatters the CPU), or exactly --hnsw-rows."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import synth
    from oracle.cport import exact_topk_c
    from oracle.hnsw import HnswIndex, num_threads

    max_rows = min(args.rows, args.hnsw_rows) if args.hnsw_rows else args.rows
    budget = float("inf") if args.hnsw_rows else args.hnsw_build_budget
    block = 25_000
    corpus = np.empty((max_rows, args.dim), np.float32)
    ix = HnswIndex(args.dim, max_rows)
    index_rows, build_s, blk = 0, 0.0, 0
    while index_rows < max_rows and build_s < budget:
        m = min(block, max_rows - index_rows)
        # iid rows: one seeded block at a time, so that generation stays out of the build clock
        corpus[index_rows:index_rows + m] = synth.make_corpus(m, args.dim, seed=SEED + blk, ties=False)
        t0 = time.perf_counter()
        ix.add(corpus[index_rows:index_rows + m])
        build_s += time.perf_counter() - t0
        index_rows += m
        blk += 1
    corpus = corpus[:index_rows]
    nq = max(16, args.hnsw_queries)
    q, _ = synth.make_queries(corpus[:block], nq, seed=7, tie_probe=False)
    for _ in range(max(1, args.warmup)):
        ix.search(q[:16], args.k)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ids, _ = ix.search(q, args.k)
    dt = time.perf_counter() - t0
    qps = nq * args.steps / dt
    nref = min(64, nq)
    ref, _, _ = exact_topk_c(corpus, q[:nref], args.k)
    recall = float(np.mean([len(set(ids[b]) & set(ref[b])) / args.k for b in range(nref)]))
    ix.close()
    cover = ("the full corpus" if index_rows == args.rows else
             f"the part of the {args.rows}-row corpus that could be inserted within the build budget "
             f"({args.hnsw_build_budget:.0f} s); HNSW cost grows ~log N, so this flatters the CPU"
             if not args.hnsw_rows else f"a subsample of the {args.rows}-row corpus; this flatters the CPU")
    sample = (f"{nq} queries per step against an HNSW index over {index_rows} rows ({cover}), "
              f"index build {build_s:.1f} s (not timed), recall@{args.k} {recall:.3f} vs exact; {HNSW_LABEL}")
    line = {
        "impl": "reference", "metric": METRIC, "value": qps, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.rows}x{args.dim} fp32 corpus, top-{args.k} cosine, "
                               f"query batch {args.batch} per GPU (CPU arm: {nq}-query sample per step, "
                               f"index over {index_rows} rows)"},
        "cpu_baseline": {"value": qps, "unit": UNIT, "cores": num_threads(), "kind": "port", "sample": sample,
                         "recall_at_k": recall, "index_rows": index_rows, "build_s": build_s},
        "e2e": {"value": qps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def build_store(torch, args, device, rank, world):
    """Synthetic FRIDA-shaped corpus (iid Gaussian rows, L2-normalised), generated on
    the device block by block and appended through cmw_store_append_f32."""
    from cmw_rag_b200 import DenseStore

    row_shard = args.shard == "rows" and world > 1
    id_offset = rank * args.rows if row_shard else 0
    st = DenseStore(args.dim, args.rows, device=device.index, f32=not args.no_f32, bf16=True,
                    id_offset=id_offset)
    g = torch.Generator(device=device).manual_seed(SEED + (rank if row_shard else 0))
    block = 125_000
    first = None
    for lo in range(0, args.rows, block):
        m = min(block, args.rows - lo)
        x = torch.randn((m, args.dim), generator=g, device=device, dtype=torch.float32)
        x = torch.nn.functional.normalize(x, dim=1)
        st.append(x)
        if first is None:
            first = x[: min(m, 65536)].clone()
        del x
    return st, first


def make_queries(torch, first, batch, dim, device, seed):
    """75 % planted needles normalise(C[j] + 0.75 g), 25 % random unit vectors."""
    g = torch.Generator(device=device).manual_seed(seed)
    noise = torch.nn.functional.normalize(torch.randn((batch, dim), generator=g, device=device), dim=1)
    j = torch.randint(0, first.shape[0], (batch,), generator=g, device=device)
    q = first[j] + 0.75 * noise
    rnd = torch.rand((batch,), generator=g, device=device) < 0.25
    q[rnd] = noise[rnd]
    needle = torch.where(rnd, torch.full_like(j, -1), j)
    return torch.nn.functional.normalize(q, dim=1).contiguous(), needle


def run_ours(args):
    import torch
    import torch.distributed as dist

    from cmw_rag_b200 import _native as N
    from cmw_rag_b200 import merge_topk

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    device = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(device)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    row_shard = args.shard == "rows" and world > 1

    st, first = build_store(torch, args, device, rank, world)
    q_seed = 7 if row_shard else 7 + rank  # row shards answer the SAME queries; replicas different ones
    q, needle = make_queries(torch, first, args.batch, args.dim, device, q_seed)
    if row_shard:  # every shard must answer the same queries; the needles live in rank 0's shard
        dist.broadcast(q, 0)
        dist.broadcast(needle, 0)
    # the end-to-end leg hands the library page-locked host buffers (driver contract: "host->device copy
    # of that step's inputs from pinned host memory")
    from cmw_rag_b200.engine import pinned_empty

    k, B = args.k, args.batch
    q_host = pinned_empty((B, args.dim), np.float32)
    q_host[:] = q.cpu().numpy()
    out_host = (pinned_empty((B, k), np.float32), pinned_empty((B, k), np.int64), np.zeros((B,), np.int32))

    peer_ex = None
    if row_shard and args.exchange == "peer":
        from cmw_rag_b200.sharded import PeerExchange

        peer_ex = PeerExchange(device=local_rank, max_batch=B, max_k=k)

    def step_device():
        if not row_shard:
            return st.search(q, k, mode=args.mode, algo=args.algo)
        sc, ids, fl, s64 = st.search(q, k, mode=args.mode, algo=args.algo, return_scores64=True)
        if peer_ex is not None:
            ms, mi, _ = peer_ex.exchange_merge(s64, ids, k)
            return ms, mi, fl
        g_s = torch.empty((world * B, k), dtype=s64.dtype, device=device)
        g_i = torch.empty((world * B, k), dtype=ids.dtype, device=device)
        dist.all_gather_into_tensor(g_s, s64)
        dist.all_gather_into_tensor(g_i, ids)
        ms, mi, _ = merge_topk(g_s.view(world, B, k), g_i.view(world, B, k), k)
        return ms, mi, fl

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    # ---- warm-up -------------------------------------------------------------------------
    for _ in range(max(3, args.warmup)):
        out = step_device()
    st.search_host(q_host, k, mode=args.mode, algo=args.algo)
    torch.cuda.synchronize(device)
    sc0, ids0, fl0 = out
    ids0_h = ids0.cpu().numpy()
    needle_h = needle.cpu().numpy()
    planted = needle_h >= 0
    if not row_shard or rank == 0:
        top1 = ids0_h[:, 0] - (st.id_offset if not row_shard else 0)
        assert (top1[planted] == needle_h[planted]).all(), "planted needles are not top-1: wrong results"
    uncertified = int(fl0.sum().item())

    # ---- timed region: device-resident ------------------------------------------------------
    sampler = ClockSampler(local_rank)
    barrier()
    time.sleep(args.leg_gap)
    if rank == 0:
        sampler.start()
    N.profile_enable(True)
    launches0 = N.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    barrier()
    dev_ms = e0.elapsed_time(e1)
    launches = N.kernel_launches() - launches0
    prof = N.profile_read()
    N.profile_enable(False)

    # ---- timed region: end to end through the host-buffer C ABI ------------------------------
    # (a) blocking calls, one after the other (cmw_search_host): H2D, kernels, D2H, sync -- per-call latency
    # (every leg starts from an idle GPU: the step is power-capped, so a leg that runs right behind another
    # one inherits its heat and reads ~5 % lower)
    barrier()
    time.sleep(args.leg_gap)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        st.search_host(q_host, k, mode=args.mode, algo=args.algo, out=out_host)
    torch.cuda.synchronize(device)
    e2e_serial_ms = (time.perf_counter() - t0) * 1e3
    if not row_shard:
        assert (out_host[1] == ids0_h).all(), "host-buffer path and device path disagree"
    # (b) the pipelined form (cmw_search_host_submit / _wait), `--in-flight` requests outstanding, as the
    # reference's callers are (S concurrent awaits per request, concurrent requests): every step still
    # copies its own inputs from pinned host memory and reads its own results back, inside the timed
    # region; the copies of one step overlap the kernels of its neighbours
    depth = max(1, min(args.in_flight, N.HOST_SLOTS))
    outs = [out_host] + [(pinned_empty((B, k), np.float32), pinned_empty((B, k), np.int64), np.zeros((B,), np.int32))
                         for _ in range(depth - 1)]
    for o in outs:
        o[1][:] = -7
    # first use of a slot allocates its staging buffers and workspace: keep that out of the clock
    for t in [st.search_host_submit(q_host, k, mode=args.mode, algo=args.algo, out=outs[i]) for i in range(depth)]:
        st.search_host_wait(t)
    from collections import deque

    pending = deque()
    barrier()
    time.sleep(args.leg_gap)
    t0 = time.perf_counter()
    for i in range(args.steps):
        if len(pending) == depth:
            st.search_host_wait(pending.popleft())
        pending.append(st.search_host_submit(q_host, k, mode=args.mode, algo=args.algo, out=outs[i % depth]))
    while pending:
        st.search_host_wait(pending.popleft())
    torch.cuda.synchronize(device)
    e2e_ms = (time.perf_counter() - t0) * 1e3
    if not row_shard:
        for o in outs[: min(depth, args.steps)]:
            assert (o[1] == ids0_h).all(), "pipelined host-buffer path and device path disagree"
    clocks = sampler.stop() if rank == 0 else None

    if world > 1:
        t = torch.tensor([dev_ms, e2e_ms, e2e_serial_ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_ms, e2e_serial_ms = float(t[0]), float(t[1]), float(t[2])

    # ---- batch-1 legs (HBM-bound regime) ------------------------------------------------------
    # "auto": the default exact path (K2 tensor-core filter over the bf16 tiles + fp64 rescoring);
    # "fp32_scan": CMW_ALGO_SCAN, K1 streaming the fp32 tiles -- BASELINE.json config 2 read literally
    # ("1M x 1536 fp32 corpus, batch 1").  Both return the oracle's ids.
    def batch1_leg(algo, elt_bytes, kernel_name):
        q1 = q[:1].contiguous()
        for _ in range(5):
            st.search(q1, k, mode=args.mode, algo=algo)
        torch.cuda.synchronize(device)
        iters = 50
        # latency: per-query CUDA events, phase profiling OFF (its event pairs cost a few microseconds per query)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
        ev[0].record()
        for i in range(iters):
            st.search(q1, k, mode=args.mode, algo=algo)
            ev[i + 1].record()
        torch.cuda.synchronize(device)
        lat = np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(iters)])
        # kernel time of the filter phase: a second loop with the library's phase timers on
        N.profile_enable(True)
        for i in range(iters):
            st.search(q1, k, mode=args.mode, algo=algo)
        torch.cuda.synchronize(device)
        prof1 = N.profile_read()
        N.profile_enable(False)
        bytes_scan = args.rows * (args.dim * elt_bytes + 4)
        filt_ms = prof1["filter"][0] / iters
        pk = measured_peaks()
        t_host0 = time.perf_counter()
        for _ in range(20):
            st.search_host(q_host[:1], k, mode=args.mode, algo=algo)
        host_ms = (time.perf_counter() - t_host0) / 20 * 1e3
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                tr = json.load(f).get("gemm_topk_kernel_batch1" if algo == "auto" else "scan_kernel_batch1")
            if tr and (tr["rows"], tr["dim"], tr["k"]) == (args.rows, args.dim, k):
                traffic = tr["traffic_bytes_per_step"]
        except (OSError, ValueError, KeyError):
            pass
        return {
            "algo": algo, "qps": 1e3 / float(np.mean(lat)), "p50_ms": float(np.median(lat)),
            "p99_ms": float(np.percentile(lat, 99)), "e2e_qps": 1e3 / host_ms, "e2e_ms": host_ms,
            "roofline": {"bound": "hbm", "kernel": kernel_name, "algorithmic_bytes": bytes_scan,
                         "achieved": bytes_scan / (filt_ms * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                         "frac": bytes_scan / (filt_ms * 1e-3) / 1e9 / pk["hbm_gbs"], "traffic": traffic,
                         "peak_source": pk["source"], "filter_ms": filt_ms,
                         "whole_query_frac": bytes_scan / (float(np.mean(lat)) * 1e-3) / 1e9 / pk["hbm_gbs"]},
        }

    b1 = None
    b1_scan = None
    if not args.skip_b1 and not row_shard and rank == 0:
        gemm_ok = st.info()["gemm_ready"] and int(N.get_option("scan_max_batch")) < 1
        if gemm_ok:
            b1 = batch1_leg("auto", 2, "gemm_topk_kernel (K2, NT=16: streams the bf16 tiles)")
        if not args.no_f32:
            b1_scan = batch1_leg("scan", 4 if args.mode == "f32" else 2, "scan_kernel (K1)")
        if b1 is None:
            b1 = b1_scan
    if world > 1:
        dist.barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- report -------------------------------------------------------------------------------
    peaks = measured_peaks()
    units = B * args.steps * (1 if row_shard else world)
    value = units / (dev_ms * 1e-3)
    e2e = units / (e2e_ms * 1e-3)
    total_rows = args.rows * (world if row_shard else 1)
    filt_ms, filt_n = prof["filter"]
    info = st.info()
    used_gemm = info["gemm_ready"] and B > int(N.get_option("scan_max_batch")) and not args.algo.startswith("scan")
    if used_gemm:
        flops = 2.0 * B * args.rows * args.dim * args.steps
        ach = flops / (filt_ms * 1e-3) / 1e12
        # B200_PROFILING.md: burst cuBLAS figure for a kernel timed in a short region, the sustained
        # (power-capped, seconds-long) one for a long step
        long_region = dev_ms > 2000.0
        pk = peaks["bf16_tflops_sustained"] if long_region else peaks["bf16_tflops"]
        roof = {"bound": "tensor", "kernel": "gemm_topk_kernel (K2)", "achieved": ach,
                "peak": pk, "unit": "TFLOP/s", "frac": ach / pk, "traffic": None,
                "peak_source": peaks["source"] + (" (sustained cuBLAS bf16: timed region > 2 s)" if long_region else
                                                  f" (burst cuBLAS bf16: timed region {dev_ms:.0f} ms)"),
                "frac_of_burst": ach / peaks["bf16_tflops"],
                "frac_of_sustained": ach / peaks["bf16_tflops_sustained"]}
    else:
        elt = 2 if args.mode == "bf16" else 4
        passes = (B + 1) // 2
        nbytes = float(passes) * args.rows * (args.dim * elt + 4) * args.steps
        ach = nbytes / (filt_ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": "scan_kernel (K1)", "achieved": ach, "peak": peaks["hbm_gbs"],
                "unit": "GB/s", "frac": ach / peaks["hbm_gbs"], "traffic": None, "peak_source": peaks["source"]}
    # DRAM traffic of the dominant kernel from the committed `ncu --set full` capture (per step = the
    # sum over that kernel's launches in one step), when this run is the workload that was profiled
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            tr = json.load(f).get("gemm_topk_kernel" if used_gemm else "scan_kernel")
        if tr and (tr["rows"], tr["dim"], tr["batch"], tr["k"]) == (args.rows, args.dim, B, k):
            roof["traffic"] = tr["traffic_bytes_per_step"]
            roof["traffic_unit"] = "bytes per step (all launches of the kernel in one step)"
            roof["algorithmic_hbm_bytes_per_step"] = tr["algorithmic_hbm_bytes_per_step"]
            roof["traffic_source"] = tr["source"]
    except (OSError, ValueError, KeyError):
        pass
    roof["kernel_ms_per_step"] = filt_ms / args.steps
    roof["kernel_launches_per_step"] = filt_n / args.steps
    roof["share_of_step"] = filt_ms / dev_ms
    phases = {name: {"ms_per_step": ms / args.steps, "launches_per_step": n / args.steps}
              for name, (ms, n) in prof.items()}

    # bf16 mode (approximate) on the same batch: throughput and recall@k against the exact mode's ids
    bf16_leg = None
    if args.mode == "f32" and not row_shard:
        for _ in range(2):
            st.search(q, k, mode="bf16", algo=args.algo)
        torch.cuda.synchronize(device)
        time.sleep(args.leg_gap)
        b0, b1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record()
        for _ in range(args.steps):
            sc_b, ids_b, _ = st.search(q, k, mode="bf16", algo=args.algo)
        b1e.record()
        torch.cuda.synchronize(device)
        ids_bh = ids_b.cpu().numpy()
        rec = float(np.mean([len(np.intersect1d(ids_bh[i], ids0_h[i])) / k for i in range(B)]))
        err = float((sc_b - sc0).abs().max().item())
        bf16_leg = {"qps": B * args.steps / (b0.elapsed_time(b1e) * 1e-3), "recall_at_k": rec,
                    "max_abs_score_err_vs_exact": err, "tolerance": 2e-3}

    cpu = None
    parity = None

    def cpu_legs():
        nonlocal cpu, parity
        # the CPU legs run on the SAME corpus bits (read back from the HBM store) and the same queries
        index_rows = min(args.rows, args.hnsw_rows or 50_000)
        if args.no_f32 or args.skip_cpu_exact:
            corpus = host_corpus(index_rows, args.dim)
            import synth

            cq, _ = synth.make_queries(corpus, max(16, args.hnsw_queries), seed=7, tie_probe=False)
        else:
            corpus = st.read_rows(0, args.rows)[0]
            cq = np.ascontiguousarray(q_host[: max(16, args.hnsw_queries)])
        qps, recall, build_s, threads = cpu_hnsw_sample(corpus, cq, k, index_rows)
        cpu = {"value": qps, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{cq.shape[0]} queries against an HNSW index over a {index_rows}-row subsample "
                         f"(build {build_s:.1f} s, not timed; HNSW cost grows ~log N, so the subsample flatters "
                         f"the CPU), recall@{k} {recall:.3f} vs exact; {HNSW_LABEL}",
               "recall_at_k": recall, "index_rows": index_rows, "build_s": build_s}
        if not (args.no_f32 or args.skip_cpu_exact):
            from oracle.cport import exact_topk_c

            nchk = min(args.cpu_queries, B)
            xq, dt, xthreads = cpu_exact_sample(corpus, cq[:nchk], k)
            cpu["exact_port"] = {"value": xq, "unit": UNIT, "cores": xthreads,
                                 "sample": f"{nchk} queries x {args.rows} rows, exact fp64 brute force "
                                           f"(oracle/c/oracle_topk.c, OpenMP, {dt:.1f} s)"}
            nsg = min(128, B)
            sq, sdt, sthreads, sids = cpu_exact_sgemm_sample(corpus, cq[:nsg], k)
            agree = float(np.mean([len(np.intersect1d(sids[i], ids0_h[i])) / k for i in range(min(nsg, nchk))]))
            cpu["exact_sgemm"] = {"value": sq, "unit": UNIT, "cores": sthreads,
                                  "sample": f"{nsg} queries x {args.rows} rows, fp32 sgemm (torch CPU) + top-{k} per "
                                            f"query ({sdt:.1f} s); recall@{k} vs the GPU's exact ids {agree:.4f}"}
        if not (args.no_f32 or args.skip_cpu_exact) and args.mode == "f32":
            # parity of the timed workload itself: the first queries of the batch against the oracle
            ref_ids, ref_sc, _ = exact_topk_c(corpus, cq[:nchk], k)
            sc0_h = sc0.cpu().numpy()
            parity = {"queries_checked": nchk, "rows": args.rows, "k": k,
                      "ids_identical_to_fp64_oracle": bool((ids0_h[:nchk] == ref_ids).all()),
                      "max_abs_score_err": float(np.abs(sc0_h[:nchk] - ref_sc).max()), "tolerance": 1e-5}
            assert parity["ids_identical_to_fp64_oracle"], "top-k ids differ from the fp64 oracle"
            assert parity["max_abs_score_err"] <= 1e-5, parity
        del corpus

    if world == 1 and not args.skip_cpu:
        try:
            cpu_legs()
        except AssertionError:
            raise  # a parity failure invalidates the line: fail loudly
        except Exception as exc:  # host-side trouble (memory, compiler) must not lose the GPU measurement
            print(f"bench.py: CPU legs failed: {exc!r}", file=sys.stderr)
            cpu = cpu or {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                          "sample": f"CPU legs failed: {exc!r}"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16 filter + f64 rescoring" if (used_gemm and args.mode == "f32") else
                 ("f32 filter + f64 rescoring" if args.mode == "f32" else "bf16"),
        "data": "synthetic",
        "config": {
            "workload": f"{total_rows}x{args.dim} {'fp32+bf16' if not args.no_f32 else 'bf16'} corpus, "
                        f"query batch {B} per GPU, top-{k} {'exact' if args.mode == 'f32' else 'bf16'} cosine",
            "parallelism": ("single GPU" if world == 1 else
                            (f"corpus row-sharded over {world} GPUs, " + ("fused NVLink peer-store exchange + merge kernels (no collective)" if args.exchange == "peer" else "NCCL all-gather + merge kernel") if row_shard
                             else f"corpus replicated, query batch sharded over {world} GPUs (no collective)")),
            "l2": "inputs larger than L2 (corpus tiles >= 3 GB per pass vs 126 MB), no flush",
            "mode": args.mode, "algo": args.algo, "uncertified_queries": uncertified,
        },
        "clocks": clocks,
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": B * args.dim * 4,
                "d2h_bytes_per_step": B * k * 12 + B * 4, "ms_per_step": e2e_ms / args.steps,
                "in_flight": depth, "api": "cmw_search_host_submit/_wait (pinned host buffers)",
                "blocking": {"value": units / (e2e_serial_ms * 1e-3), "unit": UNIT,
                             "ms_per_step": e2e_serial_ms / args.steps, "api": "cmw_search_host"}},
        "gpu_launches": int(launches),
        "roofline": roof,
        "phases": phases,
        "cpu_baseline": cpu,
        "parity": parity,
        "bf16_mode": bf16_leg,
        "batch1": b1,
        "batch1_fp32_scan": b1_scan,
        "store": {"rows": info["rows"], "hbm_bytes": info["hbm_bytes"], "gemm_ready": info["gemm_ready"]},
    }
    print(json.dumps(line), flush=True)
    st.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
