#!/usr/bin/env python
"""What the exactness certificate costs: the headline batch (1M x 1536, batch 4096, top-100) with the rigorous
residual bound (default) and with the statistical bound, on fp16 and on bf16 tiles.  One JSON line per setting."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import bench  # noqa: E402
from cmw_rag_b200 import _native as N  # noqa: E402

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
for tiles in ("f16", "bf16"):
    st, first, q, needle = bench.simple_setup(torch, 1_000_000, 1536, dev, 4096, tiles16=tiles)
    for strict in (1, 0):
        N.set_option("strict_certificate", strict)
        for _ in range(3):
            sc, ids, fl = st.search(q, 100)
        torch.cuda.synchronize()
        N.profile_enable(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            st.search(q, 100)
        e1.record()
        torch.cuda.synchronize()
        prof = N.profile_read()
        N.profile_enable(False)
        print(json.dumps({"tiles16": tiles, "certificate": "rigorous" if strict else "statistical (8 sigma, u per format)",
                          "ms_per_step": e0.elapsed_time(e1) / 10, "uncertified": int((fl != 0).sum()),
                          "phases_ms": {k: round(v[0] / 10, 3) for k, v in prof.items() if v[0]}}), flush=True)
    N.set_option("strict_certificate", 1)
    st.close()
