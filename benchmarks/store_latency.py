#!/usr/bin/env python
"""End-to-end latency of the drop-in seam: `await top_k_search_async(store, embedding, k)` on a
B200Store, as RAGRetriever.retrieve_async calls it (rag_engine/retrieval/retriever.py:179-182 of the
reference): one query (the common production path) and the S-segment asyncio.gather fan-out.
Shipped knobs: k = TOP_K_RETRIEVE = 20, S <= 4 (.env-example:97-110)."""
import argparse, asyncio, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--dim", type=int, default=1536)
    ap.add_argument("--k", type=int, default=20)
    ap.add_argument("--iters", type=int, default=300)
    args = ap.parse_args()
    import torch
    from cmw_rag_b200 import B200Store, top_k_search_async

    store = B200Store("latency", capacity=args.rows)
    g = torch.Generator(device="cuda").manual_seed(1)
    block = 100_000
    first = None
    for lo in range(0, args.rows, block):
        m = min(block, args.rows - lo)
        x = torch.nn.functional.normalize(torch.randn((m, args.dim), generator=g, device="cuda"), dim=1)
        # the host sidecar is what add() would build; the vectors go straight to HBM
        dense = store._ensure(args.dim)
        dense.append(x, torch.arange(lo, lo + m, device="cuda", dtype=torch.int32) // 8)
        if first is None:
            first = x[:4096].cpu().numpy()
        for i in range(lo, lo + m):
            store._ids.append(f"{i:012d}")
        store._docs.extend([f"chunk text {i}" for i in range(lo, lo + m)])
        store._metas.extend([{"stable_id": f"{i:012d}", "kbId": str(1000 + i // 8), "source_file": "a.md"} for i in range(lo, lo + m)])
        store._alive.extend([True] * m)
    rng = np.random.default_rng(0)
    qs = first[rng.integers(0, 4096, 64)] + 0.5 * rng.standard_normal((64, args.dim)).astype(np.float32) / np.sqrt(args.dim)
    q_lists = [q.tolist() for q in qs]  # the embedder hands over Python lists (embedder.py:143-148)

    async def run():
        out = {}
        for seg in (1, 4, 8):
            lat = []
            for it in range(args.iters + 10):
                vecs = [q_lists[(it * seg + j) % 64] for j in range(seg)]
                t0 = time.perf_counter()
                res = await asyncio.gather(*[top_k_search_async(store, v, k=args.k) for v in vecs])
                dt = time.perf_counter() - t0
                if it >= 10:
                    lat.append(dt * 1e3)
            assert all(len(r) == args.k for r in res)
            lat = np.array(lat)
            out[f"segments_{seg}"] = {"p50_ms": float(np.median(lat)), "p99_ms": float(np.percentile(lat, 99)),
                                      "mean_ms": float(lat.mean())}
        return out

    res = asyncio.run(run())
    # the raw C-ABI call for one query, for comparison
    lat = []
    for it in range(args.iters):
        t0 = time.perf_counter()
        store.dense.search_host(qs[it % 64][None, :], args.k)
        lat.append((time.perf_counter() - t0) * 1e3)
    res["cmw_search_host_1"] = {"p50_ms": float(np.median(lat)), "p99_ms": float(np.percentile(lat, 99))}
    print(json.dumps({"config": "store_latency", "rows": args.rows, "k": args.k, **res,
                      "launch_batches": store.stats["launch_batches"], "searches": store.stats["searches"]}))


if __name__ == "__main__":
    main()
