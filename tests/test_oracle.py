"""Pin the CPU oracle: known-answer tests, numpy-vs-C agreement, golden vectors from the
reference's own RAGRetriever (tests/golden/multivector_golden.json)."""
from __future__ import annotations

import json
import os

import numpy as np
import pytest

import synth
from oracle import exact_scores, exact_topk, merge_topk
from oracle.cport import exact_topk_c
from oracle.multivector import (
    extract_numeric_kbid,
    group_by_kbid,
    group_key,
    multivector_reduce,
    normalized_ranks,
    union_dedup_cap,
)


def test_kat_reference_store_test():
    """rag_engine/tests/test_storage_vector_store.py:10-24 -- the reference's only numeric
    nearest-neighbour assertion: vectors [.1,0,0],[0,.1,0]; query [.1,0,0], k=1 -> doc1."""
    c = np.array([[0.1, 0.0, 0.0], [0.0, 0.1, 0.0]], np.float32)
    q = np.array([[0.1, 0.0, 0.0]], np.float32)
    ids, sc, _ = exact_topk(c, q, 1)
    assert ids.tolist() == [[0]]
    assert abs(sc[0, 0] - 1.0) < 1e-6  # un-normalised inputs => cosine, not IP
    ids_c, sc_c, _ = exact_topk_c(c, q, 1)
    assert ids_c.tolist() == [[0]] and abs(sc_c[0, 0] - 1.0) < 1e-6
    ids_ip, sc_ip, _ = exact_topk_c(c, q, 2, metric="ip")
    assert ids_ip.tolist() == [[0, 1]] and abs(sc_ip[0, 0] - 0.01) < 1e-7


def test_extract_numeric_kbid_matches_reference_behaviour():
    # values probed against rag_engine/utils/metadata_utils.py:20-32
    assert extract_numeric_kbid("4578-toc") == "4578"
    assert extract_numeric_kbid("abc") is None
    assert extract_numeric_kbid(12) == "12"
    assert extract_numeric_kbid(None) is None
    assert extract_numeric_kbid("") is None
    assert extract_numeric_kbid("12ab34") == "12"
    assert group_key("") is None and group_key(None) is None and group_key(0) is None
    assert group_key("abc") == "abc" and group_key("77-x") == "77"


def test_numpy_and_c_oracles_agree_config1():
    """Config 1 of BASELINE.json: 10k x 1536, 64 queries, top-20 cosine."""
    c = synth.make_corpus(10000)
    q, needle = synth.make_queries(c, 64)
    i1, s1, _ = exact_topk(c, q, 20)
    i2, s2, _ = exact_topk_c(c, q, 20)
    assert (i1 == i2).all()
    assert np.abs(s1 - s2).max() <= 1e-7
    planted = needle >= 0
    assert (i1[planted, 0] == needle[planted]).all()
    # tie fixture: duplicates of row 17 at N/2 and N-1; lower id first
    assert i1[1, :3].tolist() == [17, 5000, 9999]
    # brute-force cross-check through the full matrix
    full = exact_scores(c, q[:4])
    for b in range(4):
        order = np.lexsort((np.arange(c.shape[0]), -full[b]))[:20]
        assert order.tolist() == i1[b].tolist()


def test_topk_oracle_agrees_with_independent_libraries():
    """chromadb / hnswlib (where the reference's arithmetic for this path lives: requirements.txt `chromadb==1.3.0`)
    cannot be installed offline, so the top-k oracle is pinned a second way: against two independent third-party
    implementations of the same definition -- scikit-learn's brute-force NearestNeighbors in the `cosine` metric
    (distance = 1 - cos, the very quantity hnswlib's `cosine` space and Chroma's `distances` report:
    rag_engine/storage/vector_store.py:48-51) and scipy's cdist.  Un-normalised rows, so cosine != inner product."""
    from scipy.spatial.distance import cdist
    from sklearn.neighbors import NearestNeighbors

    rng = np.random.default_rng(5)
    for n, d, b, k in ((3000, 1536, 24, 100), (777, 96, 16, 20), (50, 8, 5, 50)):
        c = (rng.standard_normal((n, d)) * rng.uniform(0.2, 3.0, size=(n, 1))).astype(np.float32)
        q = (rng.standard_normal((b, d)) * rng.uniform(0.5, 2.0, size=(b, 1))).astype(np.float32)
        ids, sc, sc64 = exact_topk(c, q, k)
        ids_c, sc_c, _ = exact_topk_c(c, q, k)
        nn = NearestNeighbors(n_neighbors=k, algorithm="brute", metric="cosine").fit(c.astype(np.float64))
        dist, nbr = nn.kneighbors(q.astype(np.float64))
        full = 1.0 - cdist(q.astype(np.float64), c.astype(np.float64), metric="cosine")
        for i in range(b):
            # the three agree on the SET and the ORDER wherever neighbouring scores differ by more than fp64 noise
            gaps = np.abs(np.diff(np.sort(full[i])[::-1][: k + 1]))
            assert gaps.min() > 1e-12, "seeded case has a near-tie; pick another seed"
            assert ids[i].tolist() == nbr[i].tolist() == ids_c[i].tolist()
            assert np.argsort(-full[i], kind="stable")[:k].tolist() == ids[i].tolist()
        assert np.abs((1.0 - dist) - sc64).max() < 1e-12
        assert np.abs(np.take_along_axis(full, ids, axis=1) - sc64).max() < 1e-12
        # inner product: against numpy's own fp64 matmul
        ids_ip, _, sc_ip = exact_topk(c, q, k, metric="ip")
        ipm = q.astype(np.float64) @ c.astype(np.float64).T
        assert (np.argsort(-ipm, axis=1, kind="stable")[:, :k] == ids_ip).all()
        assert np.abs(np.take_along_axis(ipm, ids_ip, axis=1) - sc_ip).max() < 1e-9


def test_topk_edge_cases():
    rng = np.random.default_rng(0)
    c = rng.standard_normal((37, 16)).astype(np.float32)
    q = rng.standard_normal((3, 16)).astype(np.float32)
    # k > N -> padded with (-1, -inf)
    ids, sc, _ = exact_topk(c, q, 50)
    assert (ids[:, 37:] == -1).all() and np.isneginf(sc[:, 37:]).all()
    assert sorted(ids[0, :37].tolist()) == list(range(37))
    ids_c, sc_c, _ = exact_topk_c(c, q, 50)
    assert (ids == ids_c).all()
    # tombstones
    live = np.ones(37, bool)
    live[ids[0, 0]] = False
    ids2, _, _ = exact_topk(c, q, 5, live=live)
    ids2c, _, _ = exact_topk_c(c, q, 5, live=live)
    assert ids[0, 0] not in ids2[0].tolist()
    assert (ids2 == ids2c).all()
    # zero row and zero query: cosine defined as 0
    c[5] = 0
    q[2] = 0
    s = exact_scores(c, q)
    assert (s[:, 5] == 0).all() and (s[2] == 0).all()
    ids3, sc3, _ = exact_topk_c(c, q, 37)
    assert ids3[2].tolist() == list(range(37))  # all scores 0 -> id order
    # inner product
    ids4, sc4, _ = exact_topk(c, q[:2], 7, metric="ip")
    ids4c, sc4c, _ = exact_topk_c(c, q[:2], 7, metric="ip")
    assert (ids4 == ids4c).all() and np.allclose(sc4, sc4c, atol=1e-6)
    # id offset (row shards)
    ids5, _, _ = exact_topk_c(c, q[:1], 3, id_offset=1000)
    assert (ids5 - 1000 == exact_topk_c(c, q[:1], 3)[0]).all()


def test_merge_of_shards_equals_global_topk():
    """top-k of a union of disjoint shards = merge of per-shard top-k."""
    c = synth.make_corpus(4096, 64, seed=5)
    q, _ = synth.make_queries(c, 9, seed=6)
    gi, gs, gs64 = exact_topk_c(c, q, 10)
    parts = [exact_topk_c(c[lo : lo + 1024], q, 10, id_offset=lo) for lo in range(0, 4096, 1024)]
    ids = np.stack([p[0] for p in parts])
    sc = np.stack([p[2] for p in parts])
    mi, ms = merge_topk(ids, sc, 10)
    assert (mi == gi).all()
    assert np.abs(ms - gs).max() == 0.0
    # uneven: a shard shorter than k contributes padding
    parts = [exact_topk_c(c[:5], q, 10), exact_topk_c(c[5:], q, 10, id_offset=5)]
    mi, _ = merge_topk(np.stack([p[0] for p in parts]), np.stack([p[2] for p in parts]), 10)
    assert (mi == gi).all()


def _load_golden(golden_dir):
    with open(os.path.join(golden_dir, "multivector_golden.json")) as f:
        return json.load(f)


def _kb_gid_table(kb):
    keys: dict[str, int] = {}
    gid = np.full(len(kb), -1, np.int64)
    names = []
    for i, raw in enumerate(kb):
        k = group_key(raw)
        if k is None:
            continue
        if k not in keys:
            keys[k] = len(names)
            names.append(k)
        gid[i] = keys[k]
    return gid, names


def test_multivector_matches_reference_golden(golden_dir):
    g = _load_golden(golden_dir)
    kb = g["kb"]
    gid, names = _kb_gid_table(kb)
    assert len(g["cases"]) >= 6
    for case in g["cases"]:
        p = case["params"]
        segs = case["segments"]
        assert all(s["k"] == p["top_k_retrieve"] for s in segs)  # fan-out uses top_k_retrieve
        seg_ids = [s["ids"] for s in segs]
        seg_sc = [s["scores"] for s in segs]
        cid, cfirst, cbest = union_dedup_cap(seg_ids, seg_sc, p["prl"])
        if case["rerank_input_stable_ids"] is not None:
            # A6: the exact candidate list the reference handed to its reranker
            assert [f"{i:012d}" for i in cid] == case["rerank_input_stable_ids"], case["name"]
        if p["rerank"]:
            # the fixture's reranker = own score, stable sort desc, top_k_rerank
            order = sorted(range(len(cid)), key=lambda i: cfirst[i], reverse=True)[: p["top_k_rerank"]]
            cid2 = [cid[i] for i in order]
            sc2 = [cfirst[i] for i in order]
        else:
            cid2 = cid[: p["top_k_rerank"]]
            sc2 = [0.0] * len(cid2)
        grp = group_by_kbid(cid2, sc2, gid)
        order = grp["order"]
        if p["rerank"] and p["threshold"] is not None:
            order = [gi for gi in order if grp["max"][gi] >= p["threshold"]]
        arts = case["articles"]
        assert len(order) == len(arts), case["name"]
        ranks = normalized_ranks(len(order))
        for idx, (gi, a) in enumerate(zip(order, arts)):
            assert names[grp["gid"][gi]] == a["kb_id"], case["name"]
            assert grp["max"][gi] == pytest.approx(a["rerank_score"], abs=0), case["name"]
            assert [f"{cid2[m]:012d}" for m in grp["members"][gi]] == a["matched"], case["name"]
            assert ranks[idx] == a["normalized_rank"] and idx == a["article_rank"]


def test_multivector_reduce_array_form(golden_dir):
    g = _load_golden(golden_dir)
    gid, _ = _kb_gid_table(g["kb"])
    case = next(c for c in g["cases"] if c["name"] == "multi8_dups_cap60")
    ids = np.array([[s["ids"] for s in case["segments"]]], np.int64)
    sc = np.array([[s["scores"] for s in case["segments"]]], np.float32)
    out = multivector_reduce(ids, sc, gid, prl=60, limit=0)
    assert out["cand_n"][0] == 60
    assert [f"{i:012d}" for i in out["cand_ids"][0]] == case["rerank_input_stable_ids"]
    n = out["grp_n"][0]
    assert out["grp_cnt"][0, :n].sum() == (out["cand_grp"][0] >= 0).sum()
    # best >= first, and best is attained by some occurrence
    assert (out["cand_best"][0] >= out["cand_scores"][0]).all()
    # order is a permutation, scores non-increasing along it, ties by first appearance
    order = out["grp_order"][0, :n]
    assert sorted(order.tolist()) == list(range(n))
    m = out["grp_max"][0][order]
    assert (np.diff(m) <= 0).all()
    # padded ids are ignored
    ids2 = ids.copy()
    ids2[0, 3, 10:] = -1
    out2 = multivector_reduce(ids2, sc, gid, prl=0, limit=0)
    assert (out2["cand_ids"][0, : out2["cand_n"][0]] >= 0).all()


def test_hnsw_baseline_recall_and_order():
    """The hnswlib-equivalent CPU baseline: high recall on clustered data at Chroma's defaults, results
    best-first with distance = 1 - cosine, exact on a corpus smaller than ef."""
    from oracle.hnsw import HnswIndex

    c = synth.make_clustered_corpus(6000, 64, n_centroids=64, seed=4)
    q, _ = synth.make_queries(c, 40, seed=8, tie_probe=False)
    ix = HnswIndex(64, 6000)
    assert ix.add(c[:3000]) == 3000 and ix.add(c[3000:]) == 6000
    ids, dist = ix.search(q, 10)
    ref, ref_sc, _ = exact_topk_c(c, q, 10)
    recall = np.mean([len(set(ids[b]) & set(ref[b])) / 10 for b in range(40)])
    assert recall >= 0.9, recall
    assert (np.diff(dist, axis=1) >= -1e-6).all()
    hit = ids == ref
    assert np.abs((1.0 - dist)[hit] - ref_sc[hit]).max() < 1e-5
    small = HnswIndex(64, 50)
    small.add(c[:50])
    ids2, _ = small.search(q[:5], 10)
    ref2, _, _ = exact_topk_c(c[:50], q[:5], 10)
    assert (ids2 == ref2).all()
    ix.close()
    small.close()
