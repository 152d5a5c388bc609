#!/usr/bin/env python
"""How often does a candidate pool overflow?  Clustered corpus, rows in random order and sorted by one coordinate
(a bad realistic order), batches of 16 (wide first slab + one rest launch) and 256 (4096-row first slab, geometric
slabs), with the stride permutation of the scan order on and off; device API (flags not repaired)."""
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
    for batch in (16, 256):
        for permute in (1, 0):
            N.set_option("scan_permute", permute)
            flags = 0
            for lo in range(0, 256, batch):
                sc, ids, fl = st.search(torch.from_numpy(q[lo:lo + batch]).cuda(), k)
                torch.cuda.synchronize()
                flags += int(fl.sum())
            res[f"batch={batch},scan_permute={permute}"] = {"flagged_queries_of_256": flags}
    N.set_option("scan_permute", 1)
    out[name] = res
    st.close()
print(json.dumps(out))
