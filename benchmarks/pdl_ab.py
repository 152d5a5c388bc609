"""A/B of the programmatic dependent launch of the small-batch kernel chain (option "pdl") on one box: interleaved
runs, CUDA-event latency per search."""
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

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
for rows, k in ((100_000, 20), (1_000_000, 100)):
    st, first, q, _ = bench.simple_setup(torch, rows, 1536, dev, 32)
    for b in (1, 4, 16, 32):
        qb = q[:b].contiguous()
        res = {0: [], 1: []}
        for rep in range(6):
            for pdl in (0, 1):
                N.set_option("pdl", pdl)
                for _ in range(5):
                    st.search(qb, k)
                lat, enq, _ = bench.device_latency_loop(torch, lambda: st.search(qb, k), 100, dev)
                res[pdl].append(float(np.median(lat)))
        N.set_option("pdl", 1)
        print(json.dumps({"rows": rows, "k": k, "batch": b, "p50_ms_plain": round(min(res[0]), 4),
                          "p50_ms_pdl": round(min(res[1]), 4), "median_plain": round(float(np.median(res[0])), 4),
                          "median_pdl": round(float(np.median(res[1])), 4)}), flush=True)
    st.close()
