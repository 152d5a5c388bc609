"""Generate golden vectors by running the REFERENCE's own RAGRetriever (not our code).

Run in the build container only (needs /root/reference; the GPU box has no copy):

    python tests/golden/make_golden.py

What is pinned: the multi-vector orchestration of
``rag_engine/retrieval/retriever.py:113-319`` -- fan-out (one search per segment with
k = top_k_retrieve), ordered union / dedup by stable_id, pre-rerank cap, no-rerank
truncation or rerank hand-off, group by normalised kbId with max score, inclusive
threshold, stable sort, rank normalisation -- executed by the unmodified class against
a fake store that answers with the oracle's exact top-k.

Shim: a ``.env`` copied from the reference's ``.env-example`` plus the
two variables it misnames, a stub ``langchain_text_splitters`` (not installed; the
splitter only decides how many segments a text yields, which the fixture records), and a
whitespace tokenizer instead of ``cl100k_base`` (needs a download).
"""
from __future__ import annotations

import asyncio
import hashlib
import json
import os
import shutil
import sys
import tempfile
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _install_shim(workdir: str):
    shutil.copy(os.path.join(REF, ".env-example"), os.path.join(workdir, ".env"))
    os.environ.setdefault("CHROMA_HTTP_KEEPALIVE_SECS", "30")
    os.environ.setdefault("CHROMA_HTTP_MAX_CONNECTIONS", "10")
    os.chdir(workdir)
    sys.path.insert(0, REF)

    m = types.ModuleType("langchain_text_splitters")

    class RecursiveCharacterTextSplitter:
        def __init__(self, chunk_size, chunk_overlap):
            self.cs, self.co = int(chunk_size), int(chunk_overlap)

        @classmethod
        def from_tiktoken_encoder(cls, encoding_name=None, chunk_size=0, chunk_overlap=0, separators=None):
            return cls(chunk_size, chunk_overlap)

        def split_text(self, text):
            w = text.split(" ")
            step = max(1, self.cs - self.co)
            return [" ".join(w[i : i + self.cs]) for i in range(0, len(w), step)]

    m.RecursiveCharacterTextSplitter = RecursiveCharacterTextSplitter
    sys.modules["langchain_text_splitters"] = m

    import tiktoken

    class Enc:
        def encode(self, s):
            return s.split(" ") if s else []

        def decode(self, ids):
            return " ".join(ids)

    tiktoken.get_encoding = lambda name: Enc()


def text_vector(text: str, d: int) -> np.ndarray:
    """Deterministic unit vector for a text (stands in for the FRIDA forward pass)."""
    seed = int.from_bytes(hashlib.sha1(text.encode()).digest()[:8], "little")
    g = np.random.Generator(np.random.PCG64(seed)).standard_normal(d).astype(np.float32)
    return (g / np.linalg.norm(g)).astype(np.float32)


class Doc:
    def __init__(self, row, score, meta, text):
        self.page_content = text
        self.metadata = meta
        self.row = row
        self.score = score


def main():
    work = tempfile.mkdtemp(prefix="golden_")
    _install_shim(work)
    import synth
    from oracle import exact_topk

    from rag_engine.config.settings import settings
    from rag_engine.retrieval.retriever import RAGRetriever

    n, d = 3000, 48
    corpus = synth.make_corpus(n, d, seed=99)
    kb, _ = synth.make_kbids(n, seed=5)
    # make sure the suffix / empty / non-numeric cases are hit often in a tiny corpus
    rng = np.random.Generator(np.random.PCG64(3))
    for i in rng.integers(0, n, 300):
        kb[i] = kb[i] + "-toc" if kb[i] and not kb[i].endswith("-toc") else kb[i]
    for i in rng.integers(0, n, 150):
        kb[i] = ""
    for i in rng.integers(0, n, 150):
        kb[i] = "art" + kb[i] if kb[i] else kb[i]
    art_file = os.path.join(work, "article.md")
    with open(art_file, "w") as f:
        f.write("---\ntitle: t\n---\nbody")

    calls = []

    class FakeStore:
        async def similarity_search_async(self, query_embedding, k=5):
            q = np.asarray(query_embedding, np.float32)[None, :]
            ids, sc, _ = exact_topk(corpus, q, k)
            calls.append({"k": int(k), "ids": ids[0].tolist(), "scores": [float(x) for x in sc[0]],
                          # the embedding the reference handed to the store: lets the GPU box replay the case
                          # through the CUDA backend without importing the reference
                          "query": [float(x) for x in q[0]]})
            return [
                Doc(int(r), float(s), {"stable_id": f"{int(r):012d}", "kbId": kb[int(r)], "source_file": art_file}, f"chunk {r}")
                for r, s in zip(ids[0], sc[0])
            ]

    class FakeEmbedder:
        def embed_query(self, text):
            return text_vector(text, d).tolist()

    class FakeReranker:
        """Scores = the candidate's own (first-seen) cosine score; top_k best, stable."""

        def __init__(self):
            self.seen = None

        def rerank(self, query, scored_candidates, top_k, metadata_boost_weights=None):
            self.seen = [doc.metadata["stable_id"] for doc, _ in scored_candidates]
            out = [(doc, float(doc.score)) for doc, _ in scored_candidates]
            out.sort(key=lambda t: t[1], reverse=True)
            return out[:top_k]

    words = [f"w{i}" for i in range(4000)]

    def long_query(n_words, start):
        return " ".join(words[start : start + n_words])

    cases = []

    def run_case(name, query, *, rerank, top_k_retrieve, top_k_rerank, max_segments, seg_tokens, overlap, prl, threshold, dup_segments=False):
        settings.retrieval_multiquery_enabled = True
        settings.retrieval_multiquery_max_segments = max_segments
        settings.retrieval_multiquery_segment_tokens = seg_tokens
        settings.retrieval_multiquery_segment_overlap = overlap
        settings.retrieval_multiquery_pre_rerank_limit = prl
        settings.retrieval_query_decomp_enabled = False
        settings.rerank_score_threshold = threshold
        emb = FakeEmbedder()
        if dup_segments:
            # every second segment embeds to the same vector as its predecessor: forces
            # cross-segment duplicates so the dedup path is exercised hard
            real = emb.embed_query
            state = {"i": 0, "last": None}

            def embed_query(text):
                i = state["i"]
                state["i"] += 1
                if i % 2 == 1 and state["last"] is not None:
                    return state["last"]
                state["last"] = real(text)
                return state["last"]

            emb.embed_query = embed_query
        r = RAGRetriever(
            embedder=emb, vector_store=FakeStore(), llm_manager=None,
            top_k_retrieve=top_k_retrieve, top_k_rerank=top_k_rerank, rerank_enabled=False,
        )
        rr = None
        if rerank:
            rr = FakeReranker()
            r.reranker = rr
        calls.clear()
        arts = asyncio.run(r.retrieve_async(query))
        cases.append(
            {
                "name": name,
                "params": {
                    "rerank": rerank, "top_k_retrieve": top_k_retrieve, "top_k_rerank": top_k_rerank,
                    "max_segments": max_segments, "segment_tokens": seg_tokens, "overlap": overlap,
                    "prl": prl, "threshold": threshold,
                },
                "segments": [dict(c) for c in calls],
                "rerank_input_stable_ids": rr.seen if rr else None,
                "articles": [
                    {
                        "kb_id": a.kb_id,
                        "rerank_score": a.metadata["rerank_score"],
                        "normalized_rank": a.metadata["normalized_rank"],
                        "article_rank": a.metadata["article_rank"],
                        "matched": [c.metadata["stable_id"] for c in a.matched_chunks],
                    }
                    for a in arts
                ],
            }
        )

    run_case("single_short_norerank", long_query(20, 0), rerank=False, top_k_retrieve=20, top_k_rerank=10,
             max_segments=4, seg_tokens=448, overlap=64, prl=60, threshold=0.5)
    run_case("multi4_norerank", long_query(1500, 100), rerank=False, top_k_retrieve=20, top_k_rerank=10,
             max_segments=4, seg_tokens=448, overlap=64, prl=60, threshold=0.5)
    run_case("multi4_rerank_cap60", long_query(1700, 300), rerank=True, top_k_retrieve=20, top_k_rerank=10,
             max_segments=4, seg_tokens=448, overlap=64, prl=60, threshold=0.30)
    run_case("multi8_rerank_top50_uncapped", long_query(3500, 50), rerank=True, top_k_retrieve=50, top_k_rerank=400,
             max_segments=8, seg_tokens=448, overlap=64, prl=0, threshold=0.0)
    run_case("multi8_dups_cap60", long_query(3500, 400), rerank=True, top_k_retrieve=50, top_k_rerank=60,
             max_segments=8, seg_tokens=448, overlap=64, prl=60, threshold=0.35, dup_segments=True)
    run_case("multi8_dups_norerank", long_query(3300, 200), rerank=False, top_k_retrieve=50, top_k_rerank=25,
             max_segments=8, seg_tokens=448, overlap=64, prl=400, threshold=None, dup_segments=True)
    run_case("multi3_threshold_all_filtered", long_query(1000, 10), rerank=True, top_k_retrieve=20, top_k_rerank=10,
             max_segments=4, seg_tokens=448, overlap=64, prl=60, threshold=0.99)

    out = {
        "about": "Produced by the reference's RAGRetriever.retrieve_async (unmodified) -- see make_golden.py",
        "kb": kb,
        "n": n,
        "cases": cases,
    }
    with open(os.path.join(HERE, "multivector_golden.json"), "w") as f:
        json.dump(out, f)
    for c in cases:
        print(c["name"], "segments", len(c["segments"]), "articles", len(c["articles"]),
              "rerank_in", None if c["rerank_input_stable_ids"] is None else len(c["rerank_input_stable_ids"]))
    shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
