#!/usr/bin/env python
"""recall@k report on the CLUSTERED synthetic corpus (iid Gaussian rows are the worst case
for HNSW and the best case for tie-freeness): bf16 mode of this library and the hnswlib-equivalent CPU
baseline, both against the exact answer (this library's exact mode, itself checked against the C
oracle on a subset)."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=200_000)
    ap.add_argument("--dim", type=int, default=1536)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--queries", type=int, default=512)
    ap.add_argument("--centroids", type=int, default=4096)
    args = ap.parse_args()
    import torch
    import synth
    from cmw_rag_b200 import DenseStore
    from oracle.cport import exact_topk_c
    from oracle.hnsw import HnswIndex, num_threads

    c = synth.make_clustered_corpus(args.rows, args.dim, n_centroids=args.centroids)
    q, _ = synth.make_queries(c, args.queries, seed=7, tie_probe=False)
    st = DenseStore(args.dim, args.rows)
    st.append(c)
    sc, ids, fl = st.search_host(q, args.k, mode="f32")
    ref, _, _ = exact_topk_c(c, q[:32], args.k)
    assert (ids[:32] == ref).all(), "exact mode differs from the oracle"
    scb, idsb, _ = st.search_host(q, args.k, mode="bf16")
    rec_bf16 = float(np.mean([len(set(idsb[b]) & set(ids[b])) / args.k for b in range(args.queries)]))
    ix = HnswIndex(args.dim, args.rows)
    t0 = time.perf_counter(); ix.add(c); build = time.perf_counter() - t0
    ix.search(q[:8], args.k)
    t0 = time.perf_counter(); hid, _ = ix.search(q, args.k); dt = time.perf_counter() - t0
    rec_hnsw = float(np.mean([len(set(hid[b]) & set(ids[b])) / args.k for b in range(args.queries)]))
    print(json.dumps({"config": "recall_clustered", "rows": args.rows, "dim": args.dim, "k": args.k,
                      "queries": args.queries, "centroids": args.centroids, "uncertified_after_fallback": int(fl.sum()),
                      "recall_bf16_mode": rec_bf16, "recall_hnsw_cpu": rec_hnsw, "hnsw_qps": args.queries / dt,
                      "hnsw_build_s": build, "hnsw_threads": num_threads(),
                      "max_abs_score_err_bf16": float(np.abs(scb - sc)[idsb == ids].max())}))


if __name__ == "__main__":
    main()
