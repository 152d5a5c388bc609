#!/usr/bin/env python
"""Latency / throughput sweep over batch sizes (BASELINE.json config 5: batch 1-256 over 10M x 1536,
p50/p99 per-query latency vs the HBM roofline) and the multi-vector config 3 (512 long queries x 8
segments, top-50, union + kbId groups over 1M chunks).

    python benchmarks/sweep.py --rows 10000000 --batches 1,2,4,8,16,32,64,128,256 --iters 200
    python benchmarks/sweep.py --rows 1000000 --multivector

One JSON object per line on stdout (also appended to --out).
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--dim", type=int, default=1536)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--batches", default="1,2,4,8,16,32,64,128,256")
    ap.add_argument("--modes", default="f32,bf16")
    ap.add_argument("--algos", default="auto")
    ap.add_argument("--iters", type=int, default=100)
    ap.add_argument("--no-f32", action="store_true")
    ap.add_argument("--tiles16", default="f16", choices=["f16", "bf16"])
    ap.add_argument("--multivector", action="store_true")
    ap.add_argument("--out", default="")
    ap.add_argument("--set", action="append", default=[], metavar="NAME=VALUE",
                    help="cmw_set_option before the sweep (repeatable), e.g. --set gemm_prefetch=0")
    args = ap.parse_args()

    import torch

    import bench
    from cmw_rag_b200 import _native as N

    device = torch.device("cuda:0")
    torch.cuda.set_device(device)
    options = {}
    for kv in args.set:
        name, value = kv.split("=", 1)
        N.set_option(name, float(value))
        options[name] = float(value)

    st, first = bench.build_store(torch, args.dim, device, 0, args.rows, f32=not args.no_f32, tiles16=args.tiles16)
    peaks = bench.measured_peaks()
    out_f = open(args.out, "a") if args.out else None

    def emit(obj):
        line = json.dumps(obj)
        print(line, flush=True)
        if out_f:
            out_f.write(line + "\n")
            out_f.flush()

    if args.multivector:
        # config 3: 512 long queries x 8 segments, top-50 per segment, union (cap 60 as shipped and
        # uncapped) + kbId groups; articles of geometric length (mean 8 chunks)
        import synth

        qn, s, k = 512, 8, 50
        _, art = synth.make_kbids(args.rows)
        st2_gid = torch.from_numpy(art.astype(np.int32)).to(device)
        q, _ = bench.make_queries(torch, None, first, 0, qn * s, args.dim, device, 11, 0, 1)
        seg = q.view(qn, s, args.dim)
        for prl in (60, 0):
            for mode in args.modes.split(","):
                if mode == "f32" and args.no_f32:
                    continue

                def step():
                    sc, ids, fl = st.search(seg.reshape(qn * s, args.dim), k, mode=mode)
                    return st.multivector(ids.view(qn, s, k), sc.view(qn, s, k), prl=prl, kb_gid=st2_gid)

                for _ in range(3):
                    res = step()
                torch.cuda.synchronize()
                N.profile_enable(True)
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.iters // 10 + 2)]
                ev[0].record()
                for i in range(len(ev) - 1):
                    res = step()
                    ev[i + 1].record()
                torch.cuda.synchronize()
                prof = N.profile_read()
                N.profile_enable(False)
                ms = np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(len(ev) - 1)])
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                sc, ids, fl = st.search(seg.reshape(qn * s, args.dim), k, mode=mode)
                e0.record()
                for _ in range(20):
                    st.multivector(ids.view(qn, s, k), sc.view(qn, s, k), prl=prl, kb_gid=st2_gid)
                e1.record()
                torch.cuda.synchronize()
                emit({"config": "multivector", "rows": args.rows, "long_queries": qn, "segments": s, "k": k,
                      "prl": prl, "mode": mode, "ms_p50": float(np.median(ms)), "ms_p99": float(np.percentile(ms, 99)),
                      "long_queries_per_s": qn / (float(np.median(ms)) * 1e-3),
                      "segment_vectors_per_s": qn * s / (float(np.median(ms)) * 1e-3),
                      "k4_ms": e0.elapsed_time(e1) / 20,
                      "mean_candidates": float(res.cand_n.float().mean()), "mean_groups": float(res.grp_n.float().mean()),
                      "filter_ms": prof["filter"][0] / (len(ev) - 1)})
        return

    for mode in args.modes.split(","):
        if mode == "f32" and args.no_f32:
            continue
        for algo in args.algos.split(","):
            for b in [int(x) for x in args.batches.split(",")]:
                q, needle = bench.make_queries(torch, None, first, 0, b, args.dim, device, 100 + b, 0, 1)
                for _ in range(3):
                    sc, ids, fl = st.search(q, args.k, mode=mode, algo=algo)
                torch.cuda.synchronize()
                nd = needle.cpu().numpy()
                ok = bool(((ids[:, 0].cpu().numpy() == nd) | (nd < 0)).all())
                iters = max(10, args.iters // max(1, b // 64))
                # latency loop with the phase profiler off (its event records cost ~50 us per search), then a
                # second loop with it on for the per-phase kernel times
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
                ev[0].record()
                for i in range(iters):
                    st.search(q, args.k, mode=mode, algo=algo)
                    ev[i + 1].record()
                torch.cuda.synchronize()
                ms = np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(iters)])
                N.profile_enable(True)
                for i in range(iters):
                    st.search(q, args.k, mode=mode, algo=algo)
                torch.cuda.synchronize()
                prof = N.profile_read()
                N.profile_enable(False)
                filt = prof["filter"][0] / iters
                elt = 4 if (mode == "f32" and b <= int(N.get_option("scan_max_batch")) and algo != "gemm") else 2
                passes = (b + 1) // 2 if (b <= int(N.get_option("scan_max_batch")) and algo != "gemm") or algo == "scan" else 1
                nbytes = passes * args.rows * (args.dim * elt + 4)
                flops = 2.0 * b * args.rows * args.dim
                emit({"config": "sweep", "rows": args.rows, "batch": b, "mode": mode, "algo": algo, "k": args.k,
                      "options": options, "needles_ok": ok, "uncertified": int(fl.sum().item()),
                      "batch_ms_p50": float(np.median(ms)), "batch_ms_p99": float(np.percentile(ms, 99)),
                      "per_query_ms_p50": float(np.median(ms)) / b, "qps": b / (float(np.median(ms)) * 1e-3),
                      "filter_ms": filt, "hbm_gbs_filter": nbytes / (filt * 1e-3) / 1e9,
                      "hbm_frac_filter": nbytes / (filt * 1e-3) / 1e9 / peaks["hbm_gbs"],
                      "hbm_frac_batch": nbytes / (float(np.median(ms)) * 1e-3) / 1e9 / peaks["hbm_gbs"],
                      "tflops_filter": flops / (filt * 1e-3) / 1e12,
                      "tensor_frac_filter": flops / (filt * 1e-3) / 1e12 / peaks["bf16_tflops"],
                      "phases_ms": {kk: v[0] / iters for kk, v in prof.items()}})
    st.close()


if __name__ == "__main__":
    main()
