"""A/B of a 0/1 library option on the small-batch searches, on one box: interleaved runs, CUDA-event latency per
search.  `python benchmarks/pdl_ab.py [pdl | fused_finish]` (default pdl: the programmatic dependent launch of the
kernel chain; fused_finish: the last compaction also rescores and selects)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from cmw_rag_b200 import _native as N  # noqa: E402

OPT = sys.argv[1] if len(sys.argv) > 1 else "pdl"
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
for rows, k in ((100_000, 20), (1_000_000, 20), (1_000_000, 100)):
    st, first, q, _ = bench.simple_setup(torch, rows, 1536, dev, 32)
    for b in (1, 4, 16, 32):
        qb = q[:b].contiguous()
        res = {0: [], 1: []}
        for rep in range(6):
            for pdl in (0, 1):
                N.set_option(OPT, pdl)
                for _ in range(5):
                    st.search(qb, k)
                lat, enq, _ = bench.device_latency_loop(torch, lambda: st.search(qb, k), 100, dev)
                res[pdl].append(float(np.median(lat)))
        N.set_option(OPT, 1)
        print(json.dumps({"option": OPT, "rows": rows, "k": k, "batch": b, "p50_ms_off": round(min(res[0]), 4),
                          "p50_ms_on": round(min(res[1]), 4), "median_off": round(float(np.median(res[0])), 4),
                          "median_on": round(float(np.median(res[1])), 4)}), flush=True)
    st.close()
