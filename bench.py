#!/usr/bin/env python
"""bench.py -- queries/sec of the dense-retrieval hot path at 1M x 1536, top-100 (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the CPU arm (hnswlib-equivalent HNSW), rank 0 only

A step = one batch of B query vectors through the exact search (tensor-core filter over the 16-bit tiles, fp64
rescoring, rigorous certificate) over the HBM-resident corpus.
  N = 1   cmw_search on one GPU; `value` is device-resident throughput (CUDA events on the launching stream), `e2e`
          goes through the host-buffer C ABI (cmw_search_host_submit/_wait: pinned buffers, H2D, kernels, D2H).
  N > 1   one process per GPU (torchrun / NCCL).  The SAME 1M-row corpus is row-sharded over the ranks, every rank
          answers the same B queries against its shard, and the candidates are exchanged and merged inside the
          timed region -- the north star's multi-GPU design ("scaling": "strong"; `--shard queries` keeps round 1's
          replicated-corpus mode).  The step is the two-phase search of cmw_rag_b200/sharded.py: filter half |
          all-gather of the k-th filter scores | finish half (rescoring shared between the shards) | all-gather of
          the packed candidate blocks | merge kernel with the cross-shard certificate.
Extra records on the same line: `multivector` (config 3), `sweep` (config 5 at 1M; `--sweep-rows` for 10M),
`sustained` (seconds-long region against the sustained peak), `bf16_tiles` (the north star's literal tile format),
`config4_weak` (25M bf16 rows per GPU, batch 1024), `parity` (the timed batch against the fp64 oracle).
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
BLOCK = 125_000  # rows per generated block: block b of the corpus is the same bits on every rank and every N


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=1_000_000, help="corpus rows (in total; row-sharded when N > 1)")
    ap.add_argument("--dim", type=int, default=1536)
    ap.add_argument("--batch", type=int, default=4096, help="query vectors per step")
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--mode", default="f32", choices=["f32", "bf16"])
    ap.add_argument("--algo", default="auto")
    ap.add_argument("--tiles16", default="f16", choices=["f16", "bf16"],
                    help="format of the 16-bit tiles the tensor-core filter reads (same bytes, same tensor rate)")
    ap.add_argument("--f16-bits", type=int, default=0, help="experiment: significand bits kept in fp16 tiles (8..11)")
    ap.add_argument("--shard", default="rows", choices=["rows", "queries"],
                    help="N > 1: row-sharded corpus + exchange + merge (default), or replicated corpus, sharded batch")
    ap.add_argument("--exchange", default="auto", choices=["auto", "gather", "nccl", "peer"],
                    help="--shard rows: the two exchanges of the two-phase search over NVLink peer memory "
                         "(cmw_peer_gather: `gather`; `auto` = that when every rank can map its peers' buffers, else "
                         "NCCL), over NCCL all-gathers (`nccl`), or the one-phase search with the fused peer-memory "
                         "exchange + merge (`peer`)")
    ap.add_argument("--no-f32", action="store_true", help="16-bit tiles only")
    ap.add_argument("--cpu-queries", type=int, default=16, help="queries of the exact fp64 brute-force CPU sample")
    ap.add_argument("--parity-queries", type=int, default=0,
                    help="queries of the timed batch checked against the fp64 oracle (0 = all at N = 1, 256 at N > 1)")
    ap.add_argument("--hnsw-rows", type=int, default=0,
                    help="rows in the CPU HNSW index (bounded sample); 0 = 50k for the cpu_baseline leg of the "
                         "GPU arm (~10 s build), and for --impl reference as many rows as --hnsw-build-budget allows")
    ap.add_argument("--hnsw-build-budget", type=float, default=100.0,
                    help="--impl reference: seconds of index construction before the index is frozen")
    ap.add_argument("--hnsw-queries", type=int, default=512)
    ap.add_argument("--in-flight", type=int, default=2, help="requests outstanding in the pipelined end-to-end leg")
    ap.add_argument("--leg-gap", type=float, default=2.0, help="idle seconds before each timed leg")
    ap.add_argument("--sustained-seconds", type=float, default=3.0, help="length of the sustained leg (0 = skip)")
    ap.add_argument("--sweep-rows", type=int, default=0, help="extra latency sweep over a corpus of this many rows "
                                                               "(config 5: 10000000); skips config 4")
    ap.add_argument("--config4-rows", type=int, default=25_000_000,
                    help="rows per GPU of the config-4 weak-scaling sub-record (bf16 tiles only); 0 = skip")
    ap.add_argument("--skip-cpu-exact", action="store_true")
    ap.add_argument("--skip-b1", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-extras", action="store_true", help="only the headline legs")
    return ap.parse_args()


def host_threads() -> int:
    """Host threads this process may use.  NOT os.environ['OMP_NUM_THREADS']: torchrun sets that to 1."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:  # pragma: no cover
        return max(1, os.cpu_count() or 1)


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
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                pw.append(float(r[2]))
            except ValueError:
                continue
            for name, val in zip(names, r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        # "under load" = the samples drawing at least half of the highest power seen (the legs are separated by
        # idle gaps, which must not pull the median up to the idle clock)
        load = [c for c, w in zip(sm, pw) if pw and w >= 0.5 * max(pw)]
        return {"sm_mhz": float(np.median(load)) if load else (float(np.median(sm)) if sm else None),
                "sm_max_mhz": max(mx) if mx else None, "power_w_max": max(pw) if pw else None, "samples": len(sm),
                "samples_under_load": len(load), "sm_mhz_min": min(sm) if sm else None, "reasons": sorted(reasons)}


def measured_peaks() -> dict:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def ncu_traffic(key: str, **match):
    """DRAM bytes per step of a kernel from profiles/ncu_traffic.json (written by benchmarks/ncu_summarise.py from
    an `ncu --set full` capture), when the capture was taken on this very workload; else None."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            tr = json.load(f).get(key)
        if tr and all(tr.get(k) == v for k, v in match.items()):
            return tr
    except (OSError, ValueError):
        pass
    return None


# ------------------------------------------------------------------------------------------------
# CPU arm
# ------------------------------------------------------------------------------------------------
HNSW_LABEL = ("hnswlib-equivalent HNSW re-implementation (oracle/hnsw/hnsw_baseline.cpp), chromadb 1.3.0 "
              "defaults assumed (M=16, ef_construction=100, ef_search=100), not verifiable offline")


def host_corpus(rows: int, dim: int) -> np.ndarray:
    import synth

    return synth.make_corpus(rows, dim, seed=SEED, ties=False)


def recall_of(ids: np.ndarray, ref: np.ndarray, k: int) -> float:
    n = min(len(ids), len(ref))
    return float(np.mean([len(set(ids[b]) & set(ref[b])) / k for b in range(n)]))


def hnsw_iso_recall(dim: int, k: int, rows: int, nthreads: int, target: float = 0.9, nq: int = 256) -> dict:
    """The baseline at EQUAL answer quality: on the clustered corpus of SURVEY.md 8d (graph ANN's good case; on the
    iid corpus a beam of any practical width finds almost nothing), ef_search is doubled from Chroma's default
    until recall@k against the exact answer reaches `target`."""
    import synth
    from oracle.cport import exact_topk_c
    from oracle.hnsw import HnswIndex

    c = synth.make_clustered_corpus(rows, dim, n_centroids=max(16, rows // 50))
    q, _ = synth.make_queries(c, nq, seed=9, tie_probe=False)
    ix = HnswIndex(dim, rows)
    t0 = time.perf_counter()
    ix.add(c, nthreads)
    build_s = time.perf_counter() - t0
    ref, _, _ = exact_topk_c(c, q[:64], k, nthreads=nthreads)
    points = []
    ef = 100
    while True:
        ix.search(q[:8], k, ef, nthreads)
        t0 = time.perf_counter()
        ids, _ = ix.search(q, k, ef, nthreads)
        dt = time.perf_counter() - t0
        rec = recall_of(ids, ref, k)
        points.append({"ef_search": ef, "recall_at_k": rec, "qps": nq / dt})
        if rec >= target or ef >= 3200:
            break
        ef *= 2
    ix.close()
    return {"corpus": f"clustered, {rows} x {dim} ({max(16, rows // 50)} centroids, SURVEY 8d)", "k": k,
            "target_recall": target, "reached": points[-1]["recall_at_k"] >= target, "build_s": build_s,
            "cores": nthreads, "at_target": points[-1], "chroma_default": points[0], "points": points}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on the host cores.  The reference
    delegates the arithmetic to chromadb/hnswlib (not installable here); what is timed is the hnswlib-equivalent
    HNSW index of oracle/hnsw (Chroma's defaults, all host threads, one query per thread) answering a bounded
    sample of queries per step.  The index covers as much of the corpus as can be inserted within
    --hnsw-build-budget seconds (a full 1M x 1536 build takes ~250 s on 16 cores; HNSW query cost grows ~log N, so
    a partial index flatters the CPU), or exactly --hnsw-rows.  Thread count = the cores this process may run on
    (sched_getaffinity), NOT the OMP_NUM_THREADS=1 that torchrun exports."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import synth
    from oracle.cport import exact_topk_c
    from oracle.hnsw import HnswIndex

    nthreads = host_threads()
    try:
        import torch

        torch.set_num_threads(nthreads)
    except Exception:
        pass
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
        ix.add(corpus[index_rows:index_rows + m], nthreads)
        build_s += time.perf_counter() - t0
        index_rows += m
        blk += 1
    corpus = corpus[:index_rows]
    nq = max(16, args.hnsw_queries)
    q, _ = synth.make_queries(corpus[:block], nq, seed=7, tie_probe=False)
    for _ in range(max(1, args.warmup)):
        ix.search(q[:16], args.k, 100, nthreads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ids, _ = ix.search(q, args.k, 100, nthreads)
    dt = time.perf_counter() - t0
    qps = nq * args.steps / dt
    nref = min(64, nq)
    ref, _, _ = exact_topk_c(corpus, q[:nref], args.k, nthreads=nthreads)
    recall = recall_of(ids, ref, args.k)
    # the same index with wider beams: what recall costs on this (iid) corpus
    wider = []
    for ef in (400, 1600):
        t1 = time.perf_counter()
        ids_w, _ = ix.search(q[:128], args.k, ef, nthreads)
        wider.append({"ef_search": ef, "recall_at_k": recall_of(ids_w, ref, args.k),
                      "qps": 128 / (time.perf_counter() - t1)})
    ix.close()
    iso = None
    try:
        iso = hnsw_iso_recall(args.dim, args.k, min(100_000, args.rows), nthreads)
    except Exception as exc:  # pragma: no cover
        iso = {"error": repr(exc)}
    cover = ("the full corpus" if index_rows == args.rows else
             f"the part of the {args.rows}-row corpus that could be inserted within the build budget "
             f"({args.hnsw_build_budget:.0f} s); HNSW cost grows ~log N, so this flatters the CPU"
             if not args.hnsw_rows else f"a subsample of the {args.rows}-row corpus; this flatters the CPU")
    sample = (f"{nq} queries per step against an HNSW index over {index_rows} rows ({cover}), "
              f"index build {build_s:.1f} s (not timed), {nthreads} threads, recall@{args.k} {recall:.3f} vs exact "
              f"(the GPU arm returns the exact top-{args.k}: not the same answer quality -- see iso_recall); "
              f"{HNSW_LABEL}")
    line = {
        "impl": "reference", "metric": METRIC, "value": qps, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong" if args.gpus > 1 and args.shard == "rows" else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.rows}x{args.dim} fp32 corpus, top-{args.k} cosine, "
                               f"query batch {args.batch} (CPU arm: {nq}-query sample per step, "
                               f"index over {index_rows} rows)"},
        "cpu_baseline": {"value": qps, "unit": UNIT, "cores": nthreads, "kind": "port", "sample": sample,
                         "recall_at_k": recall, "index_rows": index_rows, "build_s": build_s,
                         "wider_beams_same_index": wider, "iso_recall": iso},
        "e2e": {"value": qps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm: data
# ------------------------------------------------------------------------------------------------
def gen_rows(torch, device, dim, lo, hi):
    """Rows [lo, hi) of the synthetic FRIDA-shaped corpus (iid Gaussian, L2-normalised), yielded
    block by block.  Block b is generated from seed SEED + b whoever asks, so every rank (and rank 0 rebuilding
    the whole corpus on the host for the oracle) sees the same bits."""
    for blk in range(lo // BLOCK, (hi + BLOCK - 1) // BLOCK):
        g = torch.Generator(device=device).manual_seed(SEED + blk)
        x = torch.randn((BLOCK, dim), generator=g, device=device, dtype=torch.float32)
        x = torch.nn.functional.normalize(x, dim=1)
        b0 = blk * BLOCK
        a, b = max(lo, b0) - b0, min(hi, b0 + BLOCK) - b0
        yield b0 + a, x[a:b]
        del x


def build_store(torch, dim, device, lo, hi, f32=True, tiles16="f16", keep_first=65536):
    """A DenseStore over rows [lo, hi) (id_offset = lo), appended through cmw_store_append_f32 with the kb_gid of
    articles of 8 chunks.  Returns (store, the first rows of the shard for planting needles)."""
    from cmw_rag_b200 import DenseStore

    st = DenseStore(dim, max(1, hi - lo), device=device.index, f32=f32, bf16=True, id_offset=lo, tiles16=tiles16)
    first = None
    for g0, x in gen_rows(torch, device, dim, lo, hi):
        gid = (torch.arange(g0, g0 + x.shape[0], device=device, dtype=torch.int64) // 8).to(torch.int32)
        st.append(x, gid)
        if first is None:
            first = x[: min(x.shape[0], keep_first)].clone()
    return st, first


def make_queries(torch, dist, first, lo, batch, dim, device, seed, rank, world):
    """75 % planted needles normalise(C[j] + 0.75 g), 25 % random unit vectors.  Query i is owned by
    rank i % world, which plants it on a row of ITS shard: every shard holds needles, and every rank ends up with
    the same batch (one all-reduce)."""
    g = torch.Generator(device=device).manual_seed(seed + 1000 * rank)
    noise = torch.nn.functional.normalize(torch.randn((batch, dim), generator=g, device=device), dim=1)
    j = torch.randint(0, first.shape[0], (batch,), generator=g, device=device)
    q = first[j] + 0.75 * noise
    rnd = torch.rand((batch,), generator=g, device=device) < 0.25
    q[rnd] = noise[rnd]
    q = torch.nn.functional.normalize(q, dim=1)
    needle = torch.where(rnd, torch.full_like(j, -1), j + lo)
    if world > 1:
        mine = (torch.arange(batch, device=device) % world) == rank
        q = torch.where(mine[:, None], q, torch.zeros_like(q))
        needle1 = torch.where(mine, needle + 1, torch.zeros_like(needle))
        dist.all_reduce(q)
        dist.all_reduce(needle1)
        needle = needle1 - 1
    return q.contiguous(), needle


def simple_setup(torch, rows, dim, device, batch, seed=7, f32=True, tiles16="f16"):
    """One-GPU helper for the scripts under benchmarks/: (store, queries [batch, dim], needle ids)."""
    st, first = build_store(torch, dim, device, 0, rows, f32=f32, tiles16=tiles16)
    q, needle = make_queries(torch, None, first, 0, batch, dim, device, seed, 0, 1)
    return st, first, q, needle


def host_corpus_from_blocks(torch, device, dim, rows) -> np.ndarray:
    out = np.empty((rows, dim), np.float32)
    for g0, x in gen_rows(torch, device, dim, 0, rows):
        out[g0:g0 + x.shape[0]] = x.cpu().numpy()
    return out


# ------------------------------------------------------------------------------------------------
# GPU arm: extra records (rank 0, N = 1)
# ------------------------------------------------------------------------------------------------
LAST_ENQUEUE_MS = 0.0  # host time per iteration of the last timed_search_loop (enqueue only, before the sync)


def timed_search_loop(torch, fn, iters, device):
    """Per-iteration CUDA-event latencies (ms) of fn() on the current stream.  The host must enqueue faster than
    the GPU drains for these to be device latencies: LAST_ENQUEUE_MS says whether it did."""
    global LAST_ENQUEUE_MS
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    for e in ev:  # create the events now (the first record() of a torch event allocates it)
        e.record()
    torch.cuda.synchronize(device)
    t0 = time.perf_counter()
    ev[0].record()
    for i in range(iters):
        fn()
        ev[i + 1].record()
    LAST_ENQUEUE_MS = (time.perf_counter() - t0) * 1e3 / max(1, iters)
    torch.cuda.synchronize(device)
    return np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(iters)])


def device_latency_loop(torch, fn, iters, device, tries=3):
    """timed_search_loop, repeated (at most `tries` times) while the HOST set the pace: per-iteration enqueue time
    within 20 % of the median latency means the events measured Python, not the GPU (seen right after large
    allocations are freed).  Returns (latencies of the last attempt, its enqueue ms, attempts)."""
    for attempt in range(1, tries + 1):
        lat = timed_search_loop(torch, fn, iters, device)
        if LAST_ENQUEUE_MS < 0.8 * float(np.median(lat)) or attempt == tries:
            return lat, LAST_ENQUEUE_MS, attempt
    return lat, LAST_ENQUEUE_MS, tries


def sweep_record(torch, st, q, k, rows, dim, device, mode, batches=(1, 2, 4, 8, 16, 32, 64, 128, 256), iters=100):
    """Config 5: batch 1-256, p50 / p99 of the batch time (and per query) against the HBM floor of one pass over
    the 16-bit tiles."""
    peaks = measured_peaks()
    floor_ms = rows * (dim * 2 + 4) / (peaks["hbm_gbs"] * 1e9) * 1e3
    out = []
    for b in batches:
        if b > q.shape[0]:
            break
        qb = q[:b].contiguous()
        for _ in range(5):
            st.search(qb, k, mode=mode)
        torch.cuda.synchronize(device)
        lat, enq, attempts = device_latency_loop(torch, lambda: st.search(qb, k, mode=mode), iters, device)
        out.append({"batch": b, "p50_ms": float(np.median(lat)), "p99_ms": float(np.percentile(lat, 99)),
                    "p50_ms_per_query": float(np.median(lat)) / b, "p99_ms_per_query": float(np.percentile(lat, 99)) / b,
                    "qps": b * 1e3 / float(np.mean(lat)), "hbm_floor_frac": floor_ms / float(np.median(lat)),
                    "host_enqueue_ms": enq, "attempts": attempts})
    return {"rows": rows, "k": k, "iters": iters, "hbm_floor_ms_per_batch": floor_ms,
            "hbm_floor": "one pass over the 16-bit tiles + row multipliers at the measured copy bandwidth "
                         "(the batch is HBM-bound up to ~250 queries)", "points": out}


def multivector_record(torch, st, q, dim, device, mode, long_queries=512, segments=8, k=50, iters=5):
    """Config 3: 512 long queries x 8 segments, top-50 per segment, union (cap 60 as shipped / uncapped) + kbId
    groups, one batched launch + K4; a sample checked against oracle.multivector_reduce."""
    from oracle.multivector import multivector_reduce

    n = long_queries * segments
    seg = q[:n].reshape(long_queries, segments, dim).contiguous()
    out = {"long_queries": long_queries, "segments": segments, "k": k, "rows": st.rows}
    gid_host = None
    for prl, name in ((60, "prl_60"), (0, "uncapped")):
        for _ in range(2):
            res, s3, i3, fl = st.search_multivector(seg, k, prl=prl, mode=mode)
        torch.cuda.synchronize(device)
        lat = timed_search_loop(torch, lambda: st.search_multivector(seg, k, prl=prl, mode=mode), iters, device)
        k4 = timed_search_loop(torch, lambda: st.multivector(i3, s3, prl=prl), iters, device)
        # parity of a sample of long queries: the kernel's reduction against the oracle's restatement of
        # retriever.py:185-242 on the SAME per-segment lists
        nchk = 32
        ids_h, sc_h = i3[:nchk].cpu().numpy(), s3[:nchk].cpu().numpy()
        if gid_host is None:
            gid_host = (np.arange(st.rows, dtype=np.int64) // 8).astype(np.int32)
        ref = multivector_reduce(ids_h, sc_h, gid_host, prl=prl)
        got = res.cpu()
        same = all(np.array_equal(getattr(got, nm).numpy()[:nchk], want) for nm, want in ref.items())
        assert same, f"multi-vector reduction differs from the oracle ({name})"
        out[name] = {"ms_per_batch": float(np.mean(lat)), "long_queries_per_s": long_queries * 1e3 / float(np.mean(lat)),
                     "k4_ms": float(np.mean(k4)), "uncertified_segments": int(fl.sum().item()),
                     "parity_long_queries_checked": nchk, "matches_oracle_multivector_reduce": bool(same),
                     "mean_candidates": float(res.cand_n.float().mean().item()),
                     "mean_groups": float(res.grp_n.float().mean().item())}
    return out


def sustained_record(torch, N, st, q, k, args, device, step_ms, rows_per_rank, local_rank):
    """The headline step repeated for `--sustained-seconds`: q/s and the K2 fraction of the SUSTAINED cuBLAS bf16
    rate once the power cap has settled (the 10-20-step headline region is a burst)."""
    steps = max(20, int(args.sustained_seconds * 1e3 / max(step_ms, 0.1)))
    sampler = ClockSampler(local_rank)
    torch.cuda.synchronize(device)
    sampler.start()
    N.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        st.search(q, k, mode=args.mode, algo=args.algo)
    e1.record()
    torch.cuda.synchronize(device)
    ms = e0.elapsed_time(e1)
    prof = N.profile_read()
    N.profile_enable(False)
    clocks = sampler.stop()
    peaks = measured_peaks()
    flops = 2.0 * q.shape[0] * rows_per_rank * args.dim * steps
    ach = flops / (prof["filter"][0] * 1e-3) / 1e12
    return {"seconds": ms * 1e-3, "steps": steps, "qps": q.shape[0] * steps / (ms * 1e-3), "ms_per_step": ms / steps,
            "k2_tflops": ach, "k2_frac_of_sustained_peak": ach / peaks["bf16_tflops_sustained"],
            "k2_frac_of_burst_peak": ach / peaks["bf16_tflops"], "k2_share_of_step": prof["filter"][0] / ms,
            "step_tflops": flops / (ms * 1e-3) / 1e12,
            "step_frac_of_sustained_peak": flops / (ms * 1e-3) / 1e12 / peaks["bf16_tflops_sustained"],
            "clocks": clocks}


def batch1_leg(torch, N, st, q, q_host, k, args, device, algo, elt_bytes, kernel_name, traffic_key):
    q1 = q[:1].contiguous()
    for _ in range(5):
        st.search(q1, k, mode=args.mode, algo=algo)
    torch.cuda.synchronize(device)
    iters = 50
    # latency: per-query CUDA events, phase profiling OFF (its event pairs cost a few microseconds per query)
    lat, enqueue_ms, _ = device_latency_loop(torch, lambda: st.search(q1, k, mode=args.mode, algo=algo), iters, device)
    # kernel time of the filter phase: a second loop with the library's phase timers on
    N.profile_enable(True)
    for _ in range(iters):
        st.search(q1, k, mode=args.mode, algo=algo)
    torch.cuda.synchronize(device)
    prof1 = N.profile_read()
    N.profile_enable(False)
    rows = st.rows
    bytes_scan = rows * (args.dim * elt_bytes + 4)
    filt_ms = prof1["filter"][0] / iters
    pk = measured_peaks()
    t_host0 = time.perf_counter()
    for _ in range(20):
        st.search_host(q_host[:1], k, mode=args.mode, algo=algo)
    host_ms = (time.perf_counter() - t_host0) / 20 * 1e3
    tr = ncu_traffic(traffic_key, rows=rows, dim=args.dim, k=k)
    return {
        "algo": algo, "qps": 1e3 / float(np.mean(lat)), "p50_ms": float(np.median(lat)),
        "p99_ms": float(np.percentile(lat, 99)), "host_enqueue_ms_per_query": enqueue_ms,
        "e2e_qps": 1e3 / host_ms, "e2e_ms": host_ms,
        "roofline": {"bound": "hbm", "kernel": kernel_name, "algorithmic_bytes": bytes_scan,
                     "achieved": bytes_scan / (filt_ms * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                     "frac": bytes_scan / (filt_ms * 1e-3) / 1e9 / pk["hbm_gbs"],
                     "traffic": tr["traffic_bytes_per_step"] if tr else None,
                     "traffic_source": tr["source"] if tr else None,
                     "peak_source": pk["source"], "filter_ms": filt_ms,
                     "filter_launches": prof1["filter"][1] / iters,
                     "whole_query_frac": bytes_scan / (float(np.mean(lat)) * 1e-3) / 1e9 / pk["hbm_gbs"]},
    }


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    from cmw_rag_b200 import _native as N
    from cmw_rag_b200.engine import pinned_empty
    from cmw_rag_b200.sharded import PeerExchange, PeerGather, ShardedSearcher, shard_bounds

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
    k, B = args.k, args.batch
    if args.f16_bits:
        N.set_option("f16_bits", args.f16_bits)

    # ---- corpus -----------------------------------------------------------------------------
    if row_shard:
        lo, hi = shard_bounds(args.rows, world)[rank]
    else:
        lo, hi = 0, args.rows
    st, first = build_store(torch, args.dim, device, lo, hi, f32=not args.no_f32, tiles16=args.tiles16)
    shard_rows = hi - lo
    q_seed = 7 if (row_shard or world == 1) else 7 + rank  # replicas answer different batches
    q, needle = make_queries(torch, dist, first, lo, B, args.dim, device, q_seed,
                             rank if row_shard else 0, world if row_shard else 1)
    q_host = pinned_empty((B, args.dim), np.float32)
    q_host[:] = q.cpu().numpy()
    out_host = (pinned_empty((B, k), np.float32), pinned_empty((B, k), np.int64), np.zeros((B,), np.int32))

    searcher = None
    peer_ex = None
    peer_gather = None
    exchange_used = args.exchange
    if row_shard:
        if args.exchange == "peer":
            peer_ex = PeerExchange(device=local_rank, max_batch=B, max_k=k)
        elif args.exchange in ("auto", "gather"):
            # one decision for all ranks: the peer path needs every rank to have mapped every peer's buffer
            ok = 1
            try:
                peer_gather = PeerGather(device=local_rank, max_batch=max(B, 1024), max_k=k)
            except Exception as exc:  # noqa: BLE001
                if args.exchange == "gather":
                    raise
                print(f"bench.py: rank {rank}: peer-memory exchange unavailable ({exc!r})", file=sys.stderr)
                ok = 0
            t_ok = torch.tensor([ok], dtype=torch.int32, device=device)
            dist.all_reduce(t_ok, op=dist.ReduceOp.MIN)
            if int(t_ok.item()) == 0:
                peer_gather = None
            exchange_used = "gather" if peer_gather is not None else "nccl"
        searcher = ShardedSearcher(st, exchange=peer_ex, gather=peer_gather)

    def step_device():
        if searcher is None:
            return st.search(q, k, mode=args.mode, algo=args.algo)
        return searcher.search(q, k, mode=args.mode, algo=args.algo)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    # ---- warm-up -------------------------------------------------------------------------
    for _ in range(max(3, args.warmup)):
        out = step_device()
    if searcher is None:
        st.search_host(q_host, k, mode=args.mode, algo=args.algo)
    else:
        searcher.search_host(q_host, k, out=out_host, mode=args.mode, algo=args.algo)
    torch.cuda.synchronize(device)
    sc0, ids0, fl0 = out[0], out[1], out[2]
    ids0_h = ids0.cpu().numpy()
    sc0_h = sc0.cpu().numpy()
    needle_h = needle.cpu().numpy()
    planted = needle_h >= 0
    # a needle in every shard must come back as top-1 (SURVEY 8e verification list)
    assert (ids0_h[planted, 0] == needle_h[planted]).all(), "planted needles are not top-1: wrong results"
    needles_per_shard = None
    if row_shard:
        needles_per_shard = [int(((needle_h >= a) & (needle_h < b)).sum()) for a, b in shard_bounds(args.rows, world)]
        assert all(n > 0 for n in needles_per_shard), needles_per_shard
    uncertified = int((fl0 != 0).sum().item())

    # Leg order: every leg starts after `--leg-gap` idle seconds, and the headline e2e leg goes FIRST -- the step is
    # power-capped (sw_power_cap at 1000 W within a second of load), so whichever leg runs later inherits the
    # earlier ones' heat and reads 3-5 % lower; `sustained` is the settled number.
    from collections import deque

    sampler = ClockSampler(local_rank)
    depth = 1
    e2e_serial_ms = None
    outs = [out_host]
    if searcher is None:
        depth = max(1, min(args.in_flight, N.HOST_SLOTS))
        outs = [out_host] + [(pinned_empty((B, k), np.float32), pinned_empty((B, k), np.int64),
                              np.zeros((B,), np.int32)) for _ in range(depth - 1)]
        # first use of a slot allocates its staging buffers and workspace: keep that out of the clock
        for t in [st.search_host_submit(q_host, k, mode=args.mode, algo=args.algo, out=outs[i]) for i in range(depth)]:
            st.search_host_wait(t)
    for o in outs:
        o[1][:] = -7
    barrier()
    time.sleep(args.leg_gap)
    if rank == 0:
        sampler.start()

    # ---- timed region 1: end to end through the public host-buffer API -----------------------------
    if searcher is None:
        # the pipelined form (cmw_search_host_submit / _wait), `--in-flight` requests outstanding, as the
        # reference's callers are (S concurrent awaits per request, concurrent requests): every step copies its
        # own inputs from pinned host memory and reads its own results back, inside the timed region; the copies
        # of one step overlap the kernels of its neighbours
        pending = deque()
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            if len(pending) == depth:
                st.search_host_wait(pending.popleft())
            pending.append(st.search_host_submit(q_host, k, mode=args.mode, algo=args.algo, out=outs[i % depth]))
        while pending:
            st.search_host_wait(pending.popleft())
        torch.cuda.synchronize(device)
        e2e_ms = (time.perf_counter() - t0) * 1e3
        for o in outs[: min(depth, args.steps)]:
            assert (o[1] == ids0_h).all(), "pipelined host-buffer path and device path disagree"
        e2e_api = "cmw_search_host_submit/_wait (pinned host buffers)"
    else:
        # row shards: the pipelined ShardedSearcher.search_host_submit/_wait on every rank, `--in-flight` batches
        # outstanding -- H2D of the (replicated) queries from pinned memory on a copy stream, the two-phase search
        # with both exchanges, D2H of the merged result on a second copy stream, every step
        depth = max(1, args.in_flight)
        outs = [out_host] + [(pinned_empty((B, k), np.float32), pinned_empty((B, k), np.int64),
                              pinned_empty((B,), np.int32)) for _ in range(depth - 1)]
        # first use of a slot allocates its query buffer and the copy streams: keep that out of the clock
        for t in [searcher.search_host_submit(q_host, k, out=outs[i], mode=args.mode, algo=args.algo)
                  for i in range(depth)]:
            searcher.search_host_wait(t)
        for o in outs:
            o[1][:] = -7
        pending = deque()
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            if len(pending) == depth:
                searcher.search_host_wait(pending.popleft())
            pending.append(searcher.search_host_submit(q_host, k, out=outs[i % depth], mode=args.mode, algo=args.algo))
        while pending:
            searcher.search_host_wait(pending.popleft())
        torch.cuda.synchronize(device)
        e2e_ms = (time.perf_counter() - t0) * 1e3
        for o in outs[: min(depth, args.steps)]:
            assert (o[1] == ids0_h).all(), "host-buffer path and device path disagree"
        e2e_api = ("ShardedSearcher.search_host_submit/_wait on every rank (pinned host buffers; each rank uploads "
                   "its 1/N of the batch, NVLink all-gather; every rank downloads the merged result)")
        # one blocking call after the other, for the per-call latency
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            searcher.search_host(q_host, k, out=out_host, mode=args.mode, algo=args.algo)
        torch.cuda.synchronize(device)
        e2e_serial_ms = (time.perf_counter() - t0) * 1e3

    # ---- timed region 2: device-resident ------------------------------------------------------
    barrier()
    time.sleep(args.leg_gap)
    N.profile_enable(True)
    if searcher is not None:
        searcher.timings = {}
    launches0 = N.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    barrier()
    dev_ms = e0.elapsed_time(e1)
    launches = N.kernel_launches() - launches0
    prof = N.profile_read()
    N.profile_enable(False)
    shard_phases = None
    if searcher is not None:
        shard_phases = searcher.phase_ms()
        searcher.timings = None

    # ---- timed region 3 (N = 1): blocking host calls, one after the other -- per-call latency ----------
    if searcher is None:
        barrier()
        time.sleep(args.leg_gap)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            st.search_host(q_host, k, mode=args.mode, algo=args.algo, out=out_host)
        torch.cuda.synchronize(device)
        e2e_serial_ms = (time.perf_counter() - t0) * 1e3
        assert (out_host[1] == ids0_h).all(), "host-buffer path and device path disagree"
    clocks = sampler.stop() if rank == 0 else None

    if world > 1:
        t = torch.tensor([dev_ms, e2e_ms, e2e_serial_ms or 0.0], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_ms = float(t[0]), float(t[1])
        e2e_serial_ms = float(t[2]) if e2e_serial_ms is not None else None
        if shard_phases is not None:  # max over ranks, phase by phase
            names = sorted(shard_phases)
            pt = torch.tensor([shard_phases[n] for n in names], dtype=torch.float64, device=device)
            dist.all_reduce(pt, op=dist.ReduceOp.MAX)
            shard_phases = {n: float(v) for n, v in zip(names, pt)}

    # ---- everything below: rank 0 only (the other ranks wait at the final barrier) -----------------
    extras = {}

    def extra(name, fn):
        """An extra record must never cost the headline line: a failure is reported in its place (parity failures
        inside a record are assertion errors and are reported the same way -- the record then carries no number)."""
        try:
            torch.cuda.synchronize(device)
            extras[name] = fn()
        except Exception as exc:  # noqa: BLE001
            print(f"bench.py: extra record `{name}` failed: {exc!r}", file=sys.stderr)
            extras[name] = {"error": repr(exc)}
            try:
                N.profile_enable(False)
            except Exception:  # pragma: no cover
                pass

    def approx_record():
        for _ in range(2):
            st.search(q, k, mode="bf16", algo=args.algo)
        torch.cuda.synchronize(device)
        time.sleep(args.leg_gap)
        lat = timed_search_loop(torch, lambda: st.search(q, k, mode="bf16", algo=args.algo), args.steps, device)
        sc_b, ids_b, _ = st.search(q, k, mode="bf16", algo=args.algo)
        ids_bh = ids_b.cpu().numpy()
        return {"tiles16": args.tiles16, "qps": B * 1e3 / float(np.mean(lat)),
                "recall_at_k": float(np.mean([len(np.intersect1d(ids_bh[i], ids0_h[i])) / k for i in range(B)])),
                "max_abs_score_err_vs_exact": float((sc_b - sc0).abs().max().item()), "tolerance": 2e-3}

    def sustained():
        time.sleep(args.leg_gap)
        return sustained_record(torch, N, st, q, k, args, device, dev_ms / args.steps, shard_rows, local_rank)

    if rank == 0 and world == 1 and not args.skip_extras:
        if not args.skip_b1:
            gemm_ok = st.info()["gemm_ready"] and int(N.get_option("scan_max_batch")) < 1
            if gemm_ok:
                extra("batch1", lambda: batch1_leg(torch, N, st, q, q_host, k, args, device, "auto", 2,
                                                   "gemm_topk_kernel (K2, NT=16: streams the 16-bit tiles)",
                                                   "gemm_topk_kernel_batch1"))
            if not args.no_f32:
                extra("batch1_fp32_scan", lambda: batch1_leg(torch, N, st, q, q_host, k, args, device, "scan",
                                                             4 if args.mode == "f32" else 2, "scan_kernel (K1)",
                                                             "scan_kernel_batch1"))
        extra("sweep", lambda: sweep_record(torch, st, q, k, args.rows, args.dim, device, args.mode))
        extra("multivector", lambda: multivector_record(torch, st, q, args.dim, device, args.mode))
        # approximate mode on the same batch: throughput and recall@k against the exact mode's ids
        if args.mode == "f32":
            extra("approx_mode", approx_record)
        # a store WITHOUT 16-bit tiles (the north star's config 2 read literally: "1M x 1536 fp32 corpus"): the
        # filter reads the fp32 rows through kind::tf32 MMAs -- one pass whatever the batch (K1: one per 4 queries)
        if not args.no_f32:
            extra("fp32_only_store", lambda: fp32_only_record(torch, N, args, device, q, ids0_h))
        if args.sustained_seconds > 0:
            extra("sustained", sustained)
        # the north star's literal tile format: the same exact search over bf16 tiles (rigorous certificate:
        # K' = 320, ~290 rows rescored per query instead of ~120), and bf16 approximate mode with recall@k
        if args.tiles16 == "f16" and not args.no_f32:
            extra("bf16_tiles", lambda: bf16_tiles_record(torch, N, args, device, q, ids0_h, sc0))

    # ---- config 4 (weak scaling): 25M bf16 rows per GPU, batch 1024 -- on every rank.  LAST of the GPU legs: it
    # allocates and frees 77 GB, and a latency leg measured right behind it read 40 % high (host-side: the Python
    # enqueue loop, not the kernels, set the pace)
    config4 = None
    want_c4 = (args.config4_rows > 0 and not args.skip_extras and args.sweep_rows == 0 and
               (world == 1 or row_shard) and args.rows == 1_000_000)
    if want_c4:
        free_b, _ = torch.cuda.mem_get_info(device)
        need_b = args.config4_rows * (args.dim * 2 + 24) + (6 << 30)
        fits = torch.tensor([1 if free_b >= need_b else 0], dtype=torch.int32, device=device)
        if world > 1:  # one decision for all ranks: the record is a collective
            dist.all_reduce(fits, op=dist.ReduceOp.MIN)
        if int(fits.item()) == 0:
            config4 = {"skipped": f"needs {need_b >> 30} GiB of HBM per GPU, {free_b >> 30} GiB free on rank {rank}"}
        elif world == 1:
            try:
                config4 = config4_record(torch, dist, N, args, device, rank, world, peer_gather)
            except Exception as exc:  # noqa: BLE001 -- an extra record must not cost the headline line
                print(f"bench.py: config4_weak failed: {exc!r}", file=sys.stderr)
                config4 = {"error": repr(exc)}
                N.profile_enable(False)
        else:
            config4 = config4_record(torch, dist, N, args, device, rank, world, peer_gather)
    if world > 1:
        dist.barrier()

    if args.sweep_rows and rank == 0 and world == 1:
        st.close()
        st = None

        def sweep_large():
            big, _ = build_store(torch, args.dim, device, 0, args.sweep_rows, f32=not args.no_f32, tiles16=args.tiles16)
            try:
                return sweep_record(torch, big, q, k, args.sweep_rows, args.dim, device, args.mode, iters=200)
            finally:
                big.close()

        extra("sweep_large", sweep_large)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- report -------------------------------------------------------------------------------
    peaks = measured_peaks()
    units = B * args.steps * (1 if (row_shard or world == 1) else world)
    value = units / (dev_ms * 1e-3)
    e2e = units / (e2e_ms * 1e-3)
    filt_ms, filt_n = prof["filter"]
    info = st.info() if st is not None else {"rows": args.rows, "hbm_bytes": 0, "gemm_ready": True}
    used_gemm = info["gemm_ready"] and B > int(N.get_option("scan_max_batch")) and not args.algo.startswith("scan")
    if used_gemm:
        flops = 2.0 * B * shard_rows * args.dim * args.steps  # per rank: this rank's rows
        ach = flops / (filt_ms * 1e-3) / 1e12
        # B200_PROFILING.md: burst cuBLAS figure for a kernel timed in a short region, the sustained
        # (power-capped, seconds-long) one for a long step
        long_region = dev_ms > 2000.0
        pk = peaks["bf16_tflops_sustained"] if long_region else peaks["bf16_tflops"]
        roof = {"bound": "tensor", "kernel": "gemm_topk_kernel_2cta (K2)" if B >= 128 else "gemm_topk_kernel (K2)",
                "achieved": ach, "peak": pk, "unit": "TFLOP/s", "frac": ach / pk, "traffic": None,
                "algorithmic_flop_per_step": flops / args.steps,
                "peak_source": peaks["source"] + (" (sustained cuBLAS bf16: timed region > 2 s)" if long_region else
                                                  f" (burst cuBLAS bf16: timed region {dev_ms:.0f} ms; see `sustained`)"),
                "frac_of_burst": ach / peaks["bf16_tflops"],
                "frac_of_sustained": ach / peaks["bf16_tflops_sustained"]}
        tr = ncu_traffic("gemm_topk_kernel", rows=shard_rows, dim=args.dim, batch=B, k=k)
    else:
        elt = 2 if args.mode == "bf16" else 4
        passes = (B + 3) // 4
        nbytes = float(passes) * shard_rows * (args.dim * elt + 4) * args.steps
        ach = nbytes / (filt_ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": "scan_kernel (K1)", "achieved": ach, "peak": peaks["hbm_gbs"],
                "unit": "GB/s", "frac": ach / peaks["hbm_gbs"], "traffic": None, "peak_source": peaks["source"]}
        tr = ncu_traffic("scan_kernel", rows=shard_rows, dim=args.dim, batch=B, k=k)
    if tr:
        roof["traffic"] = tr["traffic_bytes_per_step"]
        roof["traffic_unit"] = "bytes per step (all launches of the kernel in one step)"
        roof["algorithmic_hbm_bytes_per_step"] = tr["algorithmic_hbm_bytes_per_step"]
        roof["traffic_source"] = tr["source"]
    roof["kernel_ms_per_step"] = filt_ms / args.steps
    roof["kernel_launches_per_step"] = filt_n / args.steps
    roof["share_of_step"] = filt_ms / dev_ms
    phases = {name: {"ms_per_step": ms / args.steps, "launches_per_step": n / args.steps}
              for name, (ms, n) in prof.items()}

    # ---- CPU legs + parity of the timed batch against the fp64 oracle --------------------------
    cpu = None
    parity = None

    def cpu_legs():
        nonlocal cpu, parity
        from oracle.cport import exact_topk_c, exact_topk_prefiltered, num_threads
        from oracle.hnsw import HnswIndex

        nthreads = host_threads()
        torch.set_num_threads(nthreads)
        if args.no_f32 or args.skip_cpu_exact:
            corpus = host_corpus(min(args.rows, args.hnsw_rows or 50_000), args.dim)
            import synth

            cq, _ = synth.make_queries(corpus, max(16, args.hnsw_queries), seed=7, tie_probe=False)
        else:
            # the SAME corpus bits: read back from the HBM store (N = 1) or regenerated block by block (N > 1)
            corpus = (st.read_rows(0, args.rows)[0] if not row_shard else
                      host_corpus_from_blocks(torch, device, args.dim, args.rows))
            cq = np.ascontiguousarray(q_host[: max(16, args.hnsw_queries)])
        if world == 1:
            index_rows = min(corpus.shape[0], args.hnsw_rows or 50_000)
            sub = corpus[:index_rows]
            ix = HnswIndex(sub.shape[1], sub.shape[0])
            t0 = time.perf_counter()
            ix.add(sub, nthreads)
            build_s = time.perf_counter() - t0
            ix.search(cq[:8], k, 100, nthreads)
            t0 = time.perf_counter()
            hids, _ = ix.search(cq, k, 100, nthreads)
            dt = time.perf_counter() - t0
            ref, _, _ = exact_topk_c(sub, cq[:64], k, nthreads=nthreads)
            recall = recall_of(hids, ref, k)
            ix.close()
            cpu = {"value": cq.shape[0] / dt, "unit": UNIT, "cores": nthreads, "kind": "port",
                   "sample": f"{cq.shape[0]} queries against an HNSW index over a {index_rows}-row subsample "
                             f"(build {build_s:.1f} s, not timed; HNSW cost grows ~log N, so the subsample flatters "
                             f"the CPU), recall@{k} {recall:.3f} vs exact; {HNSW_LABEL}",
                   "recall_at_k": recall, "index_rows": index_rows, "build_s": build_s}
            if not args.skip_extras:
                try:
                    cpu["iso_recall"] = hnsw_iso_recall(args.dim, k, 50_000, nthreads)
                except Exception as exc:  # pragma: no cover
                    cpu["iso_recall"] = {"error": repr(exc)}
        if args.no_f32 or args.skip_cpu_exact or args.mode != "f32":
            return
        if world == 1:
            nchk = min(args.cpu_queries, B)
            exact_topk_c(corpus[:4096], cq[:1], k, nthreads=nthreads)  # page in / thread pool warm-up
            t0 = time.perf_counter()
            exact_topk_c(corpus, cq[:nchk], k, nthreads=nthreads)
            dt = time.perf_counter() - t0
            cpu["exact_port"] = {"value": nchk / dt, "unit": UNIT, "cores": nthreads,
                                 "sample": f"{nchk} queries x {args.rows} rows, exact fp64 brute force "
                                           f"(oracle/c/oracle_topk.c, OpenMP, {dt:.1f} s)"}
        # parity of the timed batch itself.  The oracle's exact top-k for thousands of queries at 1M rows: fp32
        # sgemm prefilter (all host threads) keeping k + 156 rows per query, margin asserted, then the oracle's fp64
        # arithmetic and order on those rows (oracle/cport.py: exact_topk_prefiltered)
        nq = args.parity_queries or (B if world == 1 else min(B, 256))
        nq = min(nq, B)
        t0 = time.perf_counter()
        ref_ids, ref_sc, _ = exact_topk_prefiltered(corpus, q_host[:nq], k, nthreads=nthreads)
        dt = time.perf_counter() - t0
        if cpu is not None:
            cpu["exact_sgemm"] = {"value": nq / dt, "unit": UNIT, "cores": nthreads,
                                  "sample": f"{nq} queries x {args.rows} rows: fp32 sgemm (torch CPU) + top-{k + 156} "
                                            f"per query + fp64 rescoring in the oracle's order ({dt:.1f} s) -- the "
                                            f"exact answer, i.e. the CPU path at the GPU arm's answer quality"}
        # spot check of the prefiltered oracle against the plain brute-force oracle
        spot = min(8, nq)
        bf_ids, _, _ = exact_topk_c(corpus, q_host[:spot], k, nthreads=nthreads)
        assert (bf_ids == ref_ids[:spot]).all(), "prefiltered oracle disagrees with the brute-force oracle"
        parity = {"queries_checked": nq, "of_batch": B, "rows": args.rows, "k": k,
                  "ids_identical_to_fp64_oracle": bool((ids0_h[:nq] == ref_ids).all()),
                  "queries_with_any_id_mismatch": int((ids0_h[:nq] != ref_ids).any(axis=1).sum()),
                  "max_abs_score_err": float(np.abs(sc0_h[:nq] - ref_sc).max()), "tolerance": 1e-5,
                  "uncertified_queries": uncertified,
                  "oracle": "oracle.cport.exact_topk_prefiltered (sgemm prefilter with asserted margin + fp64 "
                            "rescoring), spot-checked against exact_topk_c",
                  "needles_top1": int(planted.sum()), "needles_per_shard": needles_per_shard}
        assert parity["ids_identical_to_fp64_oracle"], "top-k ids differ from the fp64 oracle"
        assert parity["max_abs_score_err"] <= 1e-5, parity
        del corpus

    if not args.skip_cpu:
        try:
            cpu_legs()
        except AssertionError:
            raise  # a parity failure invalidates the line: fail loudly
        except Exception as exc:  # host-side trouble (memory, compiler) must not lose the GPU measurement
            print(f"bench.py: CPU legs failed: {exc!r}", file=sys.stderr)
            cpu = cpu or {"value": None, "unit": UNIT, "cores": host_threads(), "kind": "port",
                          "sample": f"CPU legs failed: {exc!r}"}

    filt_name = {"f16": "f16", "bf16": "bf16"}[args.tiles16]
    if row_shard:
        par = (f"corpus row-sharded over {world} GPUs ({shard_rows} rows each), same {B} queries on every rank; " +
               ("one-phase search + fused NVLink peer-store exchange + merge kernels (no collective)"
                if exchange_used == "peer" else
                "two-phase search: filter | exchange of the k-th filter scores | rescoring shared between the shards | "
                "exchange of the packed candidate blocks | merge kernel + cross-shard certificate; exchanges = " +
                ("peer stores over NVLink into every rank's gather buffer + epoch flags, read in place "
                 "(cmw_peer_gather; no NCCL on the data path)" if exchange_used == "gather" else
                 "NCCL all-gathers")))
    elif world > 1:
        par = f"corpus replicated, query batch sharded over {world} GPUs (no collective)"
    else:
        par = "single GPU"
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
        "scaling": "strong" if row_shard else "weak", "vs_baseline": None,
        "dtype": (f"{filt_name} filter + f64 rescoring" if (used_gemm and args.mode == "f32") else
                  ("f32 filter + f64 rescoring" if args.mode == "f32" else filt_name)),
        "data": "synthetic",
        "config": {
            "workload": f"{args.rows}x{args.dim} {'fp32+' if not args.no_f32 else ''}{filt_name} corpus, "
                        f"query batch {B}, top-{k} {'exact' if args.mode == 'f32' else 'approximate'} cosine",
            "parallelism": par,
            "l2": "inputs larger than L2 (16-bit tiles >= 3 GB per pass vs 126 MB at N = 1; at N > 1 a shard's tiles "
                  "are re-read from HBM every step all the same: the per-step pools and candidates, 0.3 GB, evict them), "
                  "no flush",
            "exchange": exchange_used if row_shard else None,
            "mode": args.mode, "algo": args.algo, "tiles16": args.tiles16, "f16_bits": int(N.get_option("f16_bits")),
            "uncertified_queries": uncertified,
            "certificate": "rigorous" if N.get_option("strict_certificate") else "statistical",
        },
        "clocks": clocks,
        # row shards: rank g uploads rows [g B/N, (g+1) B/N) of the batch over its own PCIe link and an NVLink
        # all-gather completes it on every GPU (each query crosses PCIe once); every rank reads the whole result back
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": B * args.dim * 4,
                "d2h_bytes_per_step": (B * k * 12 + B * 4) * (world if row_shard else 1),
                "ms_per_step": e2e_ms / args.steps, "in_flight": depth, "api": e2e_api,
                "blocking": ({"value": units / (e2e_serial_ms * 1e-3), "unit": UNIT,
                              "ms_per_step": e2e_serial_ms / args.steps,
                              "api": "cmw_search_host" if searcher is None else "ShardedSearcher.search_host"}
                             if e2e_serial_ms else None)},
        "gpu_launches": int(launches),
        "roofline": roof,
        "phases": phases,
        "cpu_baseline": cpu,
        "parity": parity,
        "store": {"rows": info["rows"], "hbm_bytes": info["hbm_bytes"], "gemm_ready": info["gemm_ready"]},
    }
    if shard_phases is not None:
        coll = shard_phases.get("gather1", 0.0) + shard_phases.get("gather2", 0.0)
        mrg = shard_phases.get("kth", 0.0) + shard_phases.get("merge", 0.0)
        line["shard_phases_ms_per_step"] = {n: v / args.steps for n, v in shard_phases.items()}
        line["collective_ms_per_step"] = coll / args.steps
        line["merge_ms_per_step"] = mrg / args.steps
        per = {n: v / args.steps for n, v in shard_phases.items()}
        limiter = max(per, key=per.get)
        line["limiter"] = {"phase": limiter, "ms_per_step": per[limiter],
                           "note": "max over ranks, CUDA events at the phase boundaries of ShardedSearcher.search; "
                                   "`filter` shrinks with the shard, `finish` is shared between the shards by the "
                                   "global cut, compaction (inside `filter`) and the exchanges do not shrink"}
    if config4 is not None:
        line["config4_weak"] = config4
    line.update(extras)
    print(json.dumps(line), flush=True)
    if st is not None:
        st.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def fp32_only_record(torch, N, args, device, q, ids_ref_h):
    from cmw_rag_b200 import DenseStore

    k = args.k
    st3 = DenseStore(args.dim, args.rows, device=device.index, f32=True, bf16=False)
    for g0, x in gen_rows(torch, device, args.dim, 0, args.rows):
        st3.append(x)
    peaks = measured_peaks()
    floor_ms = args.rows * (args.dim * 4 + 4) / (peaks["hbm_gbs"] * 1e9) * 1e3
    out = {"rows": args.rows, "filter": "kind::tf32 MMAs over the fp32 tiles (gemm_topk_kernel)",
           "hbm_floor_ms_per_batch": floor_ms, "points": []}
    for b in (1, 4, 16, 64):
        qb = q[:b].contiguous()
        for _ in range(5):
            sc, ids, fl = st3.search(qb, k, mode="f32")
        torch.cuda.synchronize(device)
        assert (ids.cpu().numpy() == ids_ref_h[:b]).all() and int(fl.sum()) == 0, "tf32 filter: wrong or flagged"
        lat = timed_search_loop(torch, lambda: st3.search(qb, k, mode="f32"), 50, device)
        scan = None
        if b <= 16:
            for _ in range(3):
                st3.search(qb, k, mode="f32", algo="scan")
            scan = float(np.median(timed_search_loop(torch, lambda: st3.search(qb, k, mode="f32", algo="scan"), 20, device)))
        out["points"].append({"batch": b, "p50_ms": float(np.median(lat)), "p99_ms": float(np.percentile(lat, 99)),
                              "hbm_floor_frac": floor_ms / float(np.median(lat)), "k1_scan_p50_ms": scan})
    out["batch16_over_batch1"] = out["points"][2]["p50_ms"] / out["points"][0]["p50_ms"]
    st3.close()
    return out


def bf16_tiles_record(torch, N, args, device, q, ids_ref_h, sc_ref):
    """BASELINE.json names bf16 tiles: the same batch over a store whose 16-bit tiles are bf16 -- exact mode (ids
    must equal the fp16-tile store's, i.e. the oracle's) and bf16 approximate mode with recall@k."""
    k, B = args.k, args.batch
    st2, _ = build_store(torch, args.dim, device, 0, args.rows, f32=True, tiles16="bf16")
    for _ in range(3):
        st2.search(q, k, mode="f32", algo=args.algo)
    torch.cuda.synchronize(device)
    time.sleep(args.leg_gap)
    N.profile_enable(True)
    lat = timed_search_loop(torch, lambda: st2.search(q, k, mode="f32", algo=args.algo), args.steps, device)
    prof = N.profile_read()
    N.profile_enable(False)
    sc, ids, fl = st2.search(q, k, mode="f32", algo=args.algo)
    same = bool((ids.cpu().numpy() == ids_ref_h).all())
    assert same, "bf16-tile store and fp16-tile store disagree in exact mode"
    for _ in range(2):
        st2.search(q, k, mode="bf16", algo=args.algo)
    lat_b = timed_search_loop(torch, lambda: st2.search(q, k, mode="bf16", algo=args.algo), args.steps, device)
    sc_b, ids_b, _ = st2.search(q, k, mode="bf16", algo=args.algo)
    ids_bh = ids_b.cpu().numpy()
    rec = float(np.mean([len(np.intersect1d(ids_bh[i], ids_ref_h[i])) / k for i in range(B)]))
    out = {"exact_mode": {"qps": B * 1e3 / float(np.mean(lat)), "ms_per_step": float(np.mean(lat)),
                          "ids_identical_to_f16_tile_store": same, "uncertified_queries": int((fl != 0).sum().item()),
                          "phases_ms_per_step": {n: ms / args.steps for n, (ms, _) in prof.items()}},
           "bf16_mode": {"qps": B * 1e3 / float(np.mean(lat_b)), "recall_at_k": rec,
                         "max_abs_score_err_vs_exact": float((sc_b - sc_ref).abs().max().item()), "tolerance": 2e-3}}
    st2.close()
    return out


def config4_record(torch, dist, N, args, device, rank, world, gather=None):
    """BASELINE.json config 4 as weak scaling (200M x 1536 bf16 = 614 GB fits 8 GPUs, not 1 or 2):
    `--config4-rows` bf16 rows PER GPU, batch 1024, top-100, approximate mode (no fp32 tiles at this size), the
    exchange + merge in the timed region.  The driver's N = 1 run of this record is the one-shard baseline."""
    from cmw_rag_b200.sharded import ShardedSearcher

    rows, B, k = args.config4_rows, 1024, args.k
    lo = rank * rows
    t0 = time.perf_counter()
    st4, first = build_store(torch, args.dim, device, lo, lo + rows, f32=False, tiles16="bf16")
    torch.cuda.synchronize(device)
    build_s = time.perf_counter() - t0
    q4, needle = make_queries(torch, dist, first, lo, B, args.dim, device, 11, rank, world)
    s4 = ShardedSearcher(st4, gather=gather) if world > 1 else None

    def step():
        if s4 is None:
            return st4.search(q4, k, mode="bf16")
        return s4.search(q4, k, mode="bf16")

    for _ in range(3):
        out = step()
    torch.cuda.synchronize(device)
    ids = out[1].cpu().numpy()
    nd = needle.cpu().numpy()
    ok = bool((ids[nd >= 0, 0] == nd[nd >= 0]).all())
    assert ok, "config 4: planted needles are not top-1"
    steps = 5
    if world > 1:
        dist.barrier()
    N.profile_enable(True)
    if s4 is not None:
        s4.timings = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize(device)
    ms = e0.elapsed_time(e1)
    prof = N.profile_read()
    N.profile_enable(False)
    ph = s4.phase_ms() if s4 is not None else None
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
    peaks = measured_peaks()
    flops = 2.0 * B * rows * args.dim * steps
    rec = {"rows_per_gpu": rows, "total_rows": rows * world, "batch": B, "k": k, "mode": "bf16 (approximate), bf16 tiles",
           "ms_per_step": ms / steps, "steps": steps, "row_scans_per_s": B * rows * world * steps / (ms * 1e-3),
           "qps": B * steps / (ms * 1e-3), "needles_top1_in_every_shard": ok, "build_s": build_s,
           "k2_tflops_per_gpu": flops / (prof["filter"][0] * 1e-3) / 1e12,
           "k2_frac_of_sustained_peak": flops / (prof["filter"][0] * 1e-3) / 1e12 / peaks["bf16_tflops_sustained"],
           "scaling": "weak: rows per GPU fixed; efficiency = ms_per_step(N = 1) / ms_per_step(N)"}
    if ph is not None:
        rec["shard_phases_ms_per_step"] = {n: v / steps for n, v in ph.items()}
    st4.close()
    if world > 1:
        # STRONG scaling where the corpus is large enough for it to mean something (SURVEY 8e recommends a corpus
        # that fits one GPU): the very 25M-row corpus of the N = 1 weak record (same seeded blocks), row-sharded
        # over the N GPUs, batch 1024.  Its one-GPU time is `config4_weak.ms_per_step` of the N = 1 line.
        del st4, s4
        slo, shi = shard_bounds_local(rows, world, rank)
        st5, first5 = build_store(torch, args.dim, device, slo, shi, f32=False, tiles16="bf16")
        q5, needle5 = make_queries(torch, dist, first5, slo, B, args.dim, device, 13, rank, world)
        s5 = ShardedSearcher(st5, gather=gather)
        for _ in range(3):
            out5 = s5.search(q5, k, mode="bf16")
        torch.cuda.synchronize(device)
        nd5 = needle5.cpu().numpy()
        ok5 = bool((out5[1].cpu().numpy()[nd5 >= 0, 0] == nd5[nd5 >= 0]).all())
        assert ok5, "config 4 (strong): planted needles are not top-1"
        steps5 = 10
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps5):
            s5.search(q5, k, mode="bf16")
        e1.record()
        torch.cuda.synchronize(device)
        t5 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
        dist.all_reduce(t5, op=dist.ReduceOp.MAX)
        ms5 = float(t5[0]) / steps5
        rec["strong"] = {"total_rows": rows, "rows_per_gpu": shi - slo, "batch": B, "k": k, "ms_per_step": ms5,
                         "qps": B * 1e3 / ms5, "needles_top1_in_every_shard": ok5,
                         "one_gpu_baseline": "config4_weak.ms_per_step of the N = 1 line (the same 25M-row corpus on "
                                             "one GPU)",
                         "scaling": "strong: total rows fixed; speed-up = ms_per_step(N = 1 weak record) / ms_per_step"}
        st5.close()
    return rec


def shard_bounds_local(rows, world, rank):
    from cmw_rag_b200.sharded import shard_bounds

    return shard_bounds(rows, world)[rank]


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
