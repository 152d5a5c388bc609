#!/usr/bin/env python
"""A small end-to-end exercise of every kernel, sized for compute-sanitizer (one tool per gpurun call):
    compute-sanitizer --tool memcheck python benchmarks/sanitize_small.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import synth
from cmw_rag_b200 import DenseStore, merge_topk
from oracle.cport import exact_topk_c

n, d, k = 9000, 256, 20
c = synth.make_corpus(n, d, seed=1)
gid = (np.arange(n) // 6).astype(np.int32)
st = DenseStore(d, n + 100)
st.append(c[:5000], gid[:5000])
st.append(c[5000:], gid[5000:])
st.tombstone([5, 77, 8999])
live = np.ones(n, bool); live[[5, 77, 8999]] = False
for batch in (1, 2, 33, 300):
    q, _ = synth.make_queries(c, batch, seed=batch, tie_probe=False)
    ref, ref_sc, _ = exact_topk_c(c, q, k, live=live)
    for algo in ("auto", "scan") if batch <= 33 else ("auto",):
        sc, ids, fl = st.search_host(q, k, mode="f32", algo=algo)
        assert (ids == ref).all() and (fl == 0).all(), (batch, algo)
    sc, ids, fl = st.search_host(q, k, mode="bf16")
qd = torch.from_numpy(synth.make_queries(c, 24, seed=9, tie_probe=False)[0]).cuda()
res, s3, i3, _ = st.search_multivector(qd.view(6, 4, d), k, prl=60)
sc, ids, fl, s64 = st.search(qd, k, return_scores64=True)
ms, mi, _ = merge_topk(torch.stack([s64, s64]), torch.stack([ids, ids + n]), k)
torch.cuda.synchronize()
rows, g, lv = st.read_rows(0, 100)
assert (rows == c[:100]).all() and lv.sum() == 98
st.close()
print("sanitize_small ok")
