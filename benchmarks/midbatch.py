"""Where a mid-size batch (the HBM / tensor-pipe crossover, 32 <= B <= 1024) spends its time: per-phase kernel
time from the library's own timers, per batch size and kernel variant, against both rooflines."""
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

rows = int(os.environ.get("ROWS", "1000000"))
dim, k = 1536, 100
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
st, first = bench.build_store(torch, dim, dev, 0, rows, f32=rows <= 2_000_000)
mode = "f32" if rows <= 2_000_000 else "bf16"
q, _ = bench.make_queries(torch, None, first, 0, 4096, dim, dev, 7, 0, 1)
pk = bench.measured_peaks()
variants = [("default", {})]
for name, opts in [("1cta", {"gemm_2cta": 0}), ("2cta_from_64", {"gemm_2cta_min_batch": 64})]:
    variants.append((name, opts))
for b in (32, 64, 96, 128, 192, 256, 384, 512, 1024):
    qb = q[:b].contiguous()
    for name, opts in variants:
        old = {o: N.get_option(o) for o in opts}
        for o, v in opts.items():
            N.set_option(o, v)
        for _ in range(5):
            st.search(qb, k, mode=mode)
        torch.cuda.synchronize()
        iters = 30
        lat = bench.timed_search_loop(torch, lambda: st.search(qb, k, mode=mode), iters, dev)
        N.profile_enable(True)
        for _ in range(iters):
            st.search(qb, k, mode=mode)
        torch.cuda.synchronize()
        prof = N.profile_read()
        N.profile_enable(False)
        for o, v in old.items():
            N.set_option(o, v)
        hbm_ms = rows * (dim * 2 + 4) / (pk["hbm_gbs"] * 1e9) * 1e3
        tc_ms = 2.0 * b * rows * dim / (pk["bf16_tflops"] * 1e12) * 1e3
        print(json.dumps({"rows": rows, "batch": b, "variant": name, "p50_ms": round(float(np.median(lat)), 4),
                          "phases_ms": {n: round(ms / iters, 4) for n, (ms, c) in prof.items() if ms},
                          "launches": {n: c / iters for n, (ms, c) in prof.items() if c},
                          "hbm_floor_ms": round(hbm_ms, 4), "tensor_floor_ms": round(tc_ms, 4),
                          "frac_of_max_floor": round(max(hbm_ms, tc_ms) / float(np.median(lat)), 3)}), flush=True)
st.close()
