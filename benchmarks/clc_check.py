#!/usr/bin/env python
"""Quick parity check of the CTA-pair K2 kernel with and without cluster-launch-control scheduling."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
import numpy as np

import synth
from cmw_rag_b200 import DenseStore
from cmw_rag_b200 import _native as N
from oracle.cport import exact_topk_c

n, d, k = 60000, 1536, 100
c = synth.make_corpus(n, d)
q, _ = synth.make_queries(c, 512)
ref_ids, ref_sc, _ = exact_topk_c(c, q, k)
st = DenseStore(d, n)
st.append(c)
for clc in (0, 1, 1):
    N.set_option("gemm_clc", clc)
    print("gemm_clc", clc, flush=True)
    sc, ids, fl = st.search_host(q, k)
    print("  ids identical:", bool((ids == ref_ids).all()), "max err", float(np.abs(sc - ref_sc).max()),
          "flags", int(fl.sum()), flush=True)
    assert (ids == ref_ids).all()
print("clc_check ok")
