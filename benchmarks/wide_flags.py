#!/usr/bin/env python
"""How often does the single rest launch behind the wide first slab overflow a pool?  Clustered corpus (rows in
random order, and sorted by cluster = the worst realistic order), batch 16, device API (flags not repaired)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
import numpy as np
import torch

import synth
from cmw_rag_b200 import DenseStore
from cmw_rag_b200 import _native as N

n, d, k = 600_000, 1536, 100
c = synth.make_clustered_corpus(n, d, n_centroids=2048)
q, _ = synth.make_queries(c, 256, seed=7, tie_probe=False)
out = {"rows": n, "queries": 256, "batch": 16, "k": k}
for name in ("random_order", "sorted_by_first_coordinate"):
    if name != "random_order":
        c = np.ascontiguousarray(c[np.argsort(c[:, 0])])
    st = DenseStore(d, n)
    st.append(c)
    res = {}
    for wide in (1, 0):
        N.set_option("wide_dense", wide)
        flags = 0
        ids_all = []
        for lo in range(0, 256, 16):
            sc, ids, fl = st.search(torch.from_numpy(q[lo:lo + 16]).cuda(), k)
            torch.cuda.synchronize()
            flags += int(fl.sum())
            ids_all.append(ids.cpu().numpy())
        res[f"wide_dense={wide}"] = {"flagged_queries": flags}
        res[f"ids_{wide}"] = np.concatenate(ids_all)
    N.set_option("wide_dense", 1)
    same = bool((res.pop("ids_1") == res.pop("ids_0")).all())
    res["same_ids_both_ways"] = same
    out[name] = res
    st.close()
print(json.dumps(out))
