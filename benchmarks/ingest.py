#!/usr/bin/env python
"""Ingest throughput of the HBM store: rows/s and GB/s of K0 for device-resident rows,
and of cmw_store_append_host_f32 for pageable and page-locked host arrays.  One JSON line."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=500_000)
    ap.add_argument("--dim", type=int, default=1536)
    args = ap.parse_args()
    import torch

    from cmw_rag_b200 import DenseStore
    from cmw_rag_b200.engine import pinned_empty

    dev = torch.device("cuda:0")
    n, d = args.rows, args.dim
    gb = n * d * 4 / 1e9
    x = torch.nn.functional.normalize(torch.randn((n, d), device=dev), dim=1)
    out = {"config": "ingest", "rows": n, "dim": d, "f32_gbytes": gb}

    def timed(fn, reps=3):
        best = 1e9
        for _ in range(reps):
            st = DenseStore(d, n)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn(st)
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
            st.close()
        return best

    t = timed(lambda st: st.append(x))
    out["device_rows"] = {"seconds": t, "rows_per_s": n / t, "gbytes_per_s_in": gb / t,
                          "hbm_gbytes_per_s": gb * 2.5 / t}  # read fp32, write fp32 + bf16
    host = x.cpu().numpy()
    t = timed(lambda st: st.append(host))
    out["pageable_host_rows"] = {"seconds": t, "rows_per_s": n / t, "gbytes_per_s_in": gb / t}
    pinned = pinned_empty((n, d), np.float32)
    pinned[:] = host
    t = timed(lambda st: st.append(pinned))
    out["pinned_host_rows"] = {"seconds": t, "rows_per_s": n / t, "gbytes_per_s_in": gb / t}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
