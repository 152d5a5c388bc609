#!/usr/bin/env python
"""End-to-end throughput of the pipelined host API (cmw_search_host_submit / _wait) against the blocking
call, by number of requests in flight.  Prints one JSON line per depth."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from collections import deque

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--dim", type=int, default=1536)
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--trace", action="store_true")
    args = ap.parse_args()
    import torch

    import bench
    from cmw_rag_b200 import _native as N
    from cmw_rag_b200.engine import pinned_empty

    device = torch.device("cuda:0")
    torch.cuda.set_device(device)
    st, first, q, _ = bench.simple_setup(torch, args.rows, args.dim, device, args.batch)
    B, k = args.batch, args.k
    q_host = pinned_empty((B, args.dim), np.float32)
    q_host[:] = q.cpu().numpy()
    outs = [(pinned_empty((B, k), np.float32), pinned_empty((B, k), np.int64), np.zeros((B,), np.int32))
            for _ in range(N.HOST_SLOTS)]
    ref = st.search_host(q_host, k)[1].copy()
    # first use of every slot allocates its buffers: keep that out of the clock
    tk = [st.search_host_submit(q_host, k, out=outs[i]) for i in range(N.HOST_SLOTS)]
    for t in tk:
        st.search_host_wait(t)
    time.sleep(1.0)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        st.search_host(q_host, k, out=outs[0])
    blocking_ms = (time.perf_counter() - t0) / args.steps * 1e3
    print(json.dumps({"api": "cmw_search_host", "ms_per_step": blocking_ms, "qps": B / blocking_ms * 1e3}), flush=True)
    for overlap, depth in [(o, d) for o in (0, 1) for d in range(1, N.HOST_SLOTS + 1)]:
        N.set_option("host_overlap", overlap)
        time.sleep(1.0)  # every configuration starts from an idle (cool) GPU
        pending = deque()
        trace = []
        t0 = time.perf_counter()
        for i in range(args.steps):
            if len(pending) == depth:
                st.search_host_wait(pending.popleft())
                trace.append(("w", round((time.perf_counter() - t0) * 1e3, 2)))
            pending.append(st.search_host_submit(q_host, k, out=outs[i % depth]))
            trace.append(("s", round((time.perf_counter() - t0) * 1e3, 2)))
        while pending:
            st.search_host_wait(pending.popleft())
            trace.append(("w", round((time.perf_counter() - t0) * 1e3, 2)))
        ms = (time.perf_counter() - t0) / args.steps * 1e3
        ok = all((o[1] == ref).all() for o in outs[:depth])
        line = {"api": "cmw_search_host_submit/_wait", "host_overlap": overlap, "in_flight": depth, "ms_per_step": ms, "qps": B / ms * 1e3,
                "results_ok": bool(ok)}
        if args.trace:
            line["trace_ms"] = trace
        print(json.dumps(line), flush=True)
    st.close()


def args_ns(a):
    import argparse as ap

    return ap.Namespace(rows=a.rows, dim=a.dim, shard="queries", no_f32=False)


if __name__ == "__main__":
    main()
