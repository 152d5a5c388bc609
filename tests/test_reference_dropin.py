"""Drop-in test against the REAL reference class (build container only: needs /root/reference).

The reference's unmodified ``RAGRetriever`` (rag_engine/retrieval/retriever.py:113-319) is driven
with a ``B200Store`` as its ``vector_store`` and must return exactly the articles the golden fixture
recorded when the same class ran over a plain fake store (tests/golden/make_golden.py).  On CPU the
store's device backend is replaced by the oracle (FakeDense) -- what is under test here is the host
layer: the seam, await coalescing of the per-segment fan-out, RetrievedDoc / metadata shapes.
"""
from __future__ import annotations

import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"

WORKER = r'''
import asyncio, json, os, sys, tempfile
ROOT = {root!r}
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np
import make_golden as mg
use_gpu = {use_gpu!r}
work = tempfile.mkdtemp(prefix="dropin_")
mg._install_shim(work)
import synth
import cmw_rag_b200.store as store_mod
if not use_gpu:
    from test_host_cpu import FakeDense
    store_mod.DenseStore = FakeDense
from rag_engine.config.settings import settings
from rag_engine.retrieval.retriever import RAGRetriever

golden = json.load(open(os.path.join(ROOT, "tests", "golden", "multivector_golden.json")))
n, d = golden["n"], 48
corpus = synth.make_corpus(n, d, seed=99)
kb = golden["kb"]
art_file = os.path.join(work, "article.md")
open(art_file, "w").write("---\ntitle: t\n---\nbody")

store = store_mod.B200Store(collection_name="golden", capacity=n)
asyncio.run(store.add_async(
    texts=[f"chunk {{r}}" for r in range(n)],
    metadatas=[{{"stable_id": f"{{r:012d}}", "kbId": kb[r], "source_file": art_file}} for r in range(n)],
    ids=[f"{{r:012d}}" for r in range(n)], embeddings=corpus))

class Emb:
    def embed_query(self, text):
        return mg.text_vector(text, d).tolist()

words = [f"w{{i}}" for i in range(4000)]
queries = {{"single_short_norerank": (20, 0), "multi4_norerank": (1500, 100)}}
checked = 0
for case in golden["cases"]:
    if case["name"] not in queries or case["params"]["rerank"]:
        continue
    p = case["params"]
    settings.retrieval_multiquery_enabled = True
    settings.retrieval_multiquery_max_segments = p["max_segments"]
    settings.retrieval_multiquery_segment_tokens = p["segment_tokens"]
    settings.retrieval_multiquery_segment_overlap = p["overlap"]
    settings.retrieval_multiquery_pre_rerank_limit = p["prl"]
    settings.retrieval_query_decomp_enabled = False
    settings.rerank_score_threshold = p["threshold"]
    nw, start = queries[case["name"]]
    r = RAGRetriever(embedder=Emb(), vector_store=store, llm_manager=None,
                     top_k_retrieve=p["top_k_retrieve"], top_k_rerank=p["top_k_rerank"], rerank_enabled=False)
    before = store.stats["launch_batches"]
    arts = asyncio.run(r.retrieve_async(" ".join(words[start:start + nw])))
    launches = store.stats["launch_batches"] - before
    assert launches == 1, ("the per-segment fan-out must coalesce into one batched search", launches)
    got = [{{"kb_id": a.kb_id, "rerank_score": a.metadata["rerank_score"], "normalized_rank": a.metadata["normalized_rank"],
            "article_rank": a.metadata["article_rank"], "matched": [c.metadata["stable_id"] for c in a.matched_chunks]}} for a in arts]
    assert got == case["articles"], (case["name"], got[:2], case["articles"][:2])
    assert store.stats["max_batch"] >= len(case["segments"])
    checked += 1
assert checked == 2
print("dropin ok", checked)
'''


def _run(use_gpu: bool):
    script = WORKER.format(root=ROOT, use_gpu=use_gpu)
    res = subprocess.run([sys.executable, "-c", script], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert "dropin ok 2" in res.stdout


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "rag_engine")), reason="reference checkout not present")
def test_reference_retriever_over_b200store_host_layer():
    _run(use_gpu=False)
