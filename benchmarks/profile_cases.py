#!/usr/bin/env python
"""Short, fixed kernel sequences for the ncu captures under profiles/ (benchmarks/scripts/profile_r02.sh):

    python benchmarks/profile_cases.py step      # 6 exact searches of batch 4096 over 1M x 1536 (fp16 tiles)
    python benchmarks/profile_cases.py b1        # 12 batch-1 searches (K2 NT=16, wide first slab)
    python benchmarks/profile_cases.py tf32      # 12 batch-16 searches on a store WITHOUT 16-bit tiles (tf32 filter)
    python benchmarks/profile_cases.py scan      # 12 batch-1 searches through K1 (fp32 scan)
    python benchmarks/profile_cases.py mid 256   # 6 searches of a crossover batch (here 256)
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import bench  # noqa: E402

case = sys.argv[1]
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
if case == "tf32":
    from cmw_rag_b200 import DenseStore

    st = DenseStore(1536, 1_000_000, f32=True, bf16=False)
    first = None
    for g0, x in bench.gen_rows(torch, dev, 1536, 0, 1_000_000):
        st.append(x)
        first = x[:65536].clone() if first is None else first
    q, _ = bench.make_queries(torch, None, first, 0, 16, 1536, dev, 7, 0, 1)
    for _ in range(12):
        st.search(q, 100)
else:
    batch = 4096 if case == "step" else (int(sys.argv[2]) if case == "mid" else 1)
    st, first, q, _ = bench.simple_setup(torch, 1_000_000, 1536, dev, batch)
    for _ in range(6 if case in ("step", "mid") else 12):
        st.search(q, 100, algo="scan" if case == "scan" else "auto")
torch.cuda.synchronize()
print("done", case)
