#!/usr/bin/env python
"""Option sweeps for cmw_search (slab growth, K') at a fixed workload; prints ms per step."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--option", default="slab_growth")
    ap.add_argument("--values", default="0,2,3,4,6")
    ap.add_argument("--iters", type=int, default=8)
    args = ap.parse_args()
    import torch
    import bench
    from cmw_rag_b200 import _native as N
    dev = torch.device("cuda:0")

    ap_tiles = os.environ.get("CMW_TILES16", "f16")
    st, first, q, needle = bench.simple_setup(torch, args.rows, 1536, dev, args.batch, tiles16=ap_tiles)
    for v in [float(x) for x in args.values.split(",")]:
        N.set_option(args.option, v)
        for _ in range(3):
            sc, ids, fl = st.search(q, args.k)
        torch.cuda.synchronize()
        N.profile_enable(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            st.search(q, args.k)
        e1.record(); torch.cuda.synchronize()
        prof = N.profile_read(); N.profile_enable(False)
        print(json.dumps({"option": args.option, "value": v, "ms_per_step": e0.elapsed_time(e1) / args.iters,
                          "uncertified": int(fl.sum()),
                          "phases": {k: [round(x[0] / args.iters, 3), x[1] // args.iters] for k, x in prof.items()}}), flush=True)


if __name__ == "__main__":
    main()
