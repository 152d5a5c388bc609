"""GPU parity tests (run with ``-m gpu`` on a B200): the CUDA path, called through the C ABI, against
the CPU oracle on the same seeded inputs.  Bars (BASELINE.json north_star): fp32-exact mode -- top-k
ids bit-identical to the fp64 oracle (ties -> lower id), scores within 1e-5; bf16 mode -- scores
within 2e-3, reported as recall@k.
"""
from __future__ import annotations

import json
import os

import numpy as np
import pytest

import synth
from oracle import exact_topk, merge_topk as oracle_merge
from oracle.cport import exact_topk_c
from oracle.multivector import group_key, multivector_reduce

pytestmark = pytest.mark.gpu

F32_TOL = 1e-5
BF16_TOL = 2e-3


@pytest.fixture(scope="module")
def torch_cuda():
    import torch

    assert torch.cuda.is_available()
    return torch


@pytest.fixture(scope="module")
def cfg1(torch_cuda):
    """Config 1 of BASELINE.json: 10k x 1536, 64 queries, top-20 cosine."""
    from cmw_rag_b200 import DenseStore

    c = synth.make_corpus(10000)
    q, needle = synth.make_queries(c, 64)
    kb, _ = synth.make_kbids(10000)
    keys: dict[str, int] = {}
    gid = np.full(10000, -1, np.int32)
    for i, raw in enumerate(kb):
        k = group_key(raw)
        if k is not None:
            gid[i] = keys.setdefault(k, len(keys))
    st = DenseStore(1536, 10000)
    st.append(c, gid)
    ref_ids, ref_sc, _ = exact_topk_c(c, q, 20)
    yield dict(store=st, c=c, q=q, needle=needle, gid=gid, ref_ids=ref_ids, ref_sc=ref_sc)
    st.close()


def _check_exact(ids, scores, ref_ids, ref_sc):
    assert (ids == ref_ids).all(), f"{(ids != ref_ids).sum()} ids differ"
    fin = np.isfinite(ref_sc)
    assert np.abs(scores[fin] - ref_sc[fin]).max() <= F32_TOL
    assert (np.isneginf(scores) == np.isneginf(ref_sc)).all()


def test_config1_exact_scan_host_api(cfg1):
    sc, ids, flags = cfg1["store"].search_host(cfg1["q"], 20, mode="f32", algo="scan")
    _check_exact(ids, sc, cfg1["ref_ids"], cfg1["ref_sc"])
    assert (flags == 0).all()
    planted = cfg1["needle"] >= 0
    assert (ids[planted, 0] == cfg1["needle"][planted]).all()
    assert ids[1, :3].tolist() == [17, 5000, 9999]  # duplicated rows: lower id first


def test_config1_exact_auto_device_api(cfg1, torch_cuda):
    torch = torch_cuda
    q = torch.from_numpy(cfg1["q"]).cuda()
    sc, ids, flags = cfg1["store"].search(q, 20, mode="f32")
    torch.cuda.synchronize()
    _check_exact(ids.cpu().numpy(), sc.cpu().numpy(), cfg1["ref_ids"], cfg1["ref_sc"])
    assert int(flags.sum()) == 0


@pytest.mark.parametrize("algo", ["scan", "auto"])
def test_config1_bf16_recall(cfg1, algo):
    sc, ids, _ = cfg1["store"].search_host(cfg1["q"], 20, mode="bf16", algo=algo)
    ref = cfg1["ref_ids"]
    recall = np.mean([len(set(ids[b]) & set(ref[b])) / 20 for b in range(ref.shape[0])])
    assert recall >= 0.97, recall
    # scores of the rows that were returned agree with the exact scores of those rows
    c, q = cfg1["c"], cfg1["q"]
    ex = np.einsum("bkd,bd->bk", c[ids].astype(np.float64), q.astype(np.float64))
    assert np.abs(sc - ex).max() <= BF16_TOL
    assert (ids[np.arange(64), 0] == ref[:, 0]).all()


@pytest.mark.parametrize("batch", [1, 2, 3, 5])
def test_ragged_batches(cfg1, batch):
    sc, ids, _ = cfg1["store"].search_host(cfg1["q"][:batch], 20, mode="f32")
    _check_exact(ids, sc, cfg1["ref_ids"][:batch], cfg1["ref_sc"][:batch])


def test_inner_product(cfg1):
    c = cfg1["c"].copy()
    rng = np.random.default_rng(3)
    c *= rng.uniform(0.5, 2.0, size=(c.shape[0], 1)).astype(np.float32)  # un-normalised rows
    from cmw_rag_b200 import DenseStore

    st = DenseStore(1536, c.shape[0])
    st.append(c)
    q = cfg1["q"][:8] * 3.0
    ref_ids, ref_sc, _ = exact_topk_c(c, q, 10, metric="ip")
    for algo in ("scan", "auto"):
        sc, ids, fl = st.search_host(q, 10, metric="ip", mode="f32", algo=algo)
        assert (ids == ref_ids).all()
        assert np.abs(sc - ref_sc).max() <= 2e-5 * 6.0  # scores are up to |q||c| = 6
    ref_ids_c, ref_sc_c, _ = exact_topk_c(c, q, 10, metric="cosine")
    sc, ids, fl = st.search_host(q, 10, metric="cosine", mode="f32")
    _check_exact(ids, sc, ref_ids_c, ref_sc_c)
    st.close()


def test_edge_cases_small_dims(torch_cuda):
    from cmw_rag_b200 import DenseStore

    rng = np.random.default_rng(0)
    c = rng.standard_normal((37, 16)).astype(np.float32)
    q = rng.standard_normal((3, 16)).astype(np.float32)
    c[5] = 0
    q[2] = 0
    st = DenseStore(16, 64)
    # empty store: nothing to return
    sc, ids, _ = st.search_host(q, 5)
    assert (ids == -1).all() and np.isneginf(sc).all()
    st.append(c)
    # k > N -> padded with (-1, -inf), like Chroma returning fewer results
    ref_ids, ref_sc, _ = exact_topk(c, q, 50)
    sc, ids, _ = st.search_host(q, 50)
    _check_exact(ids, sc, ref_ids, ref_sc)
    assert ids[2, :37].tolist() == list(range(37))  # zero query: all scores 0 -> id order
    # tombstones are never returned
    live = np.ones(37, bool)
    dead = [int(ref_ids[0, 0]), int(ref_ids[0, 1]), 36]
    live[dead] = False
    st.tombstone(dead + [dead[0]])
    assert st.live_rows == 34
    ref_ids2, ref_sc2, _ = exact_topk(c, q, 37, live=live)
    sc, ids, _ = st.search_host(q, 37)
    _check_exact(ids, sc, ref_ids2, ref_sc2)
    sc, ids, _ = st.search_host(q, 37, mode="bf16")
    assert set(ids[0].tolist()) - {-1} == set(np.flatnonzero(live).tolist())
    # append after tombstoning; id offset
    st.close()
    st = DenseStore(16, 64, id_offset=1000)
    st.append(c[:20])
    st.append(c[20:])
    ref_ids, ref_sc, _ = exact_topk(c, q, 7, id_offset=1000)
    sc, ids, _ = st.search_host(q, 7)
    _check_exact(ids, sc, ref_ids, ref_sc)
    st.close()


@pytest.mark.parametrize("n,d", [(50007, 1024), (30000, 768), (20011, 200)])
def test_multi_slab_other_dims(torch_cuda, n, d):
    """Several slabs (dense + growing sparse ones), row counts that are not tile multiples, the
    register-resident (1024), the fp32-only specialisation (768) and the generic (200) paths."""
    from cmw_rag_b200 import DenseStore

    c = synth.make_corpus(n, d, seed=123)
    q, _ = synth.make_queries(c, 6, seed=5)
    st = DenseStore(d, n + 5)
    st.append(c)
    ref_ids, ref_sc, _ = exact_topk_c(c, q, 100)
    sc, ids, fl = st.search_host(q, 100, mode="f32")
    _check_exact(ids, sc, ref_ids, ref_sc)
    assert (fl == 0).all()
    sc, ids, fl = st.search_host(q[:3], 100, mode="bf16")
    recall = np.mean([len(set(ids[b]) & set(ref_ids[b])) / 100 for b in range(3)])
    assert recall >= 0.9
    st.close()


def test_adversarial_ascending_order(torch_cuda):
    """Rows sorted by ascending score: every slab floods the pool; the overflow must be flagged or
    the answer exact -- never silently wrong."""
    from cmw_rag_b200 import DenseStore

    d, n = 64, 60000
    rng = np.random.default_rng(1)
    q = rng.standard_normal((1, d)).astype(np.float32)
    c = rng.standard_normal((n, d)).astype(np.float32)
    order = np.argsort((c @ q[0]) / np.linalg.norm(c, axis=1))
    c = np.ascontiguousarray(c[order])
    st = DenseStore(d, n)
    st.append(c)
    ref_ids, ref_sc, _ = exact_topk_c(c, q, 10)
    sc, ids, fl = st.search_host(q, 10)  # host API: falls back to the overflow-proof schedule
    assert fl[0] == 0
    _check_exact(ids, sc, ref_ids, ref_sc)
    import torch

    from cmw_rag_b200 import _native as N

    # small batches take a wide first slab (65536 rows through a scratch matrix): this whole corpus fits in
    # it, so even the device API gets the exact answer without a flag
    sc1, ids1, fl1 = st.search(torch.from_numpy(q).cuda(), 10)
    torch.cuda.synchronize()
    assert int(fl1[0]) == 0
    _check_exact(ids1.cpu().numpy(), sc1.cpu().numpy(), ref_ids, ref_sc)
    # without the wide slab K2 still scans the tiles in a stride permutation: every slab is a representative
    # sample, no pool overflows
    N.set_option("wide_dense", 0)
    try:
        sc3, ids3, fl3 = st.search(torch.from_numpy(q).cuda(), 10)
        torch.cuda.synchronize()
        assert int(fl3[0]) == 0
        _check_exact(ids3.cpu().numpy(), sc3.cpu().numpy(), ref_ids, ref_sc)
        # in storage order (scan_permute = 0, or K1) every slab floods the pool: the device API reports it, the
        # overflow-proof schedule and the host API's repair chain get the exact answer
        N.set_option("scan_permute", 0)
        _, _, fl_dev = st.search(torch.from_numpy(q).cuda(), 10)  # device API: reports, never hides
        _, _, fl_k1 = st.search(torch.from_numpy(q).cuda(), 10, algo="scan")
        sc2, ids2, fl2 = st.search(torch.from_numpy(q).cuda(), 10, algo="scan_safe")
        torch.cuda.synchronize()
        assert int(fl_dev[0]) == 1 and int(fl_k1[0]) == 1 and int(fl2[0]) == 0
        _check_exact(ids2.cpu().numpy(), sc2.cpu().numpy(), ref_ids, ref_sc)
        sc, ids, fl = st.search_host(q, 10)  # host API: falls back to the overflow-proof schedule
        assert fl[0] == 0
        _check_exact(ids, sc, ref_ids, ref_sc)
    finally:
        N.set_option("wide_dense", 1)
        N.set_option("scan_permute", 1)
    st.close()


def test_wide_first_slab_paths(torch_cuda):
    """Small batches: wide first slab (two-level selection over 16 scratch segments per query) and the rest of
    the corpus in one launch.  Ties at the cut, thousands of near matches, tombstones inside the slab, corpora
    just below / at / above one wide slab, a sorted corpus whose single rest launch overflows the pool (flagged
    on the device API, repaired by the host API) -- and the answer must not depend on the option."""
    import torch

    from cmw_rag_b200 import DenseStore
    from cmw_rag_b200 import _native as N

    d, k = 64, 50
    rng = np.random.default_rng(8)
    for n in (4097, 20000, 65536, 65537, 150000):
        c = synth.make_corpus(n, d, seed=300 + n % 97, ties=False)
        q, _ = synth.make_queries(c, 5, seed=9, tie_probe=False)
        c[100:400] = c[7]  # 300 exact duplicates near the top of query 0
        q[0] = c[7] + 0.1 * q[0]
        strong = np.arange(2100) + min(n - 2200, 9000) if n >= 12000 else np.arange(0)
        c[strong] = q[1][None, :] + 0.02 * rng.standard_normal((strong.size, d)).astype(np.float32)
        live = np.ones(n, bool)
        live[rng.integers(0, n, size=n // 50)] = False
        st = DenseStore(d, n)
        st.append(c)
        st.tombstone(np.flatnonzero(~live))
        ref_ids, ref_sc, _ = exact_topk_c(c, q, k, live=live)
        for wide in (1, 0):
            N.set_option("wide_dense", wide)
            try:
                for algo in ("auto", "scan"):
                    sc, ids, fl = st.search_host(q, k, algo=algo)
                    assert (fl == 0).all(), (n, wide, algo)
                    _check_exact(ids, sc, ref_ids, ref_sc)
            finally:
                N.set_option("wide_dense", 1)
        st.close()
    # rows sorted by ascending score, more rows than one wide slab.  K2 scans the tiles in a stride permutation, so
    # the slab is a representative sample, its threshold holds for the rest and even the device API is exact
    # without a flag; K1 scans in storage order, overflows, and the host API's repair chain settles it
    n = 200000
    q = rng.standard_normal((1, d)).astype(np.float32)
    c = rng.standard_normal((n, d)).astype(np.float32)
    c = np.ascontiguousarray(c[np.argsort((c @ q[0]) / np.linalg.norm(c, axis=1))])
    st = DenseStore(d, n)
    st.append(c)
    ref_ids, ref_sc, _ = exact_topk_c(c, q, 10)
    sc_d, ids_d, fl_dev = st.search(torch.from_numpy(q).cuda(), 10)
    _, _, fl_scan = st.search(torch.from_numpy(q).cuda(), 10, algo="scan")
    torch.cuda.synchronize()
    assert int(fl_dev[0]) == 0 and int(fl_scan[0]) == 1
    _check_exact(ids_d.cpu().numpy(), sc_d.cpu().numpy(), ref_ids, ref_sc)
    for algo in ("auto", "scan"):
        sc, ids, fl = st.search_host(q, 10, algo=algo)
        assert fl[0] == 0
        _check_exact(ids, sc, ref_ids, ref_sc)
    st.close()


def test_merge_topk_matches_oracle(cfg1, torch_cuda):
    torch = torch_cuda
    from cmw_rag_b200 import DenseStore, merge_topk

    c, q = cfg1["c"], cfg1["q"][:9]
    shards = []
    bounds = [0, 2500, 2507, 7000, 10000]
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        st = DenseStore(1536, hi - lo, id_offset=lo)
        st.append(c[lo:hi])
        shards.append(st)
    qd = torch.from_numpy(q).cuda()
    parts = [st.search(qd, 20, return_scores64=True) for st in shards]
    s64 = torch.stack([p[3] for p in parts])
    ids = torch.stack([p[1] for p in parts])
    ms, mi, _ = merge_topk(s64, ids, 20)
    torch.cuda.synchronize()
    assert (mi.cpu().numpy() == cfg1["ref_ids"][:9]).all()
    assert np.abs(ms.cpu().numpy() - cfg1["ref_sc"][:9]).max() <= F32_TOL
    oi, osc = oracle_merge(ids.cpu().numpy(), s64.cpu().numpy(), 20)
    assert (oi == mi.cpu().numpy()).all()
    for st in shards:
        st.close()


def _gid_table(kb):
    keys: dict[str, int] = {}
    gid = np.full(len(kb), -1, np.int32)
    for i, raw in enumerate(kb):
        k = group_key(raw)
        if k is not None:
            gid[i] = keys.setdefault(k, len(keys))
    return gid


def _mv_compare(res, ref):
    for name, want in ref.items():
        got = getattr(res, name).numpy()
        if want.dtype.kind == "f":
            assert np.array_equal(got, want), name
        else:
            assert (got == want).all(), name


def test_multivector_kernel_matches_oracle_and_golden(golden_dir, torch_cuda):
    torch = torch_cuda
    from cmw_rag_b200 import DenseStore

    with open(os.path.join(golden_dir, "multivector_golden.json")) as f:
        g = json.load(f)
    gid = _gid_table(g["kb"])
    st = DenseStore(8, len(gid))
    st.append(np.ones((len(gid), 8), np.float32), gid)
    for case in g["cases"]:
        ids = np.array([[s["ids"] for s in case["segments"]]], np.int64)
        sc = np.array([[s["scores"] for s in case["segments"]]], np.float32)
        for prl, limit in ((case["params"]["prl"], 0), (0, 0), (case["params"]["prl"], case["params"]["top_k_rerank"]), (7, 3)):
            ref = multivector_reduce(ids, sc, gid, prl=prl, limit=limit)
            res = st.multivector(torch.from_numpy(ids).cuda(), torch.from_numpy(sc).cuda(), prl=prl, limit=limit).cpu()
            _mv_compare(res, ref)
        if case["rerank_input_stable_ids"] is not None:
            res = st.multivector(torch.from_numpy(ids).cuda(), torch.from_numpy(sc).cuda(),
                                 prl=case["params"]["prl"]).cpu()
            n = int(res.cand_n[0])
            assert [f"{i:012d}" for i in res.cand_ids[0, :n].tolist()] == case["rerank_input_stable_ids"]
    # random ragged case: padding ids, duplicates across and inside segments, many queries
    rng = np.random.default_rng(9)
    ids = rng.integers(0, 300, size=(33, 8, 50)).astype(np.int64)
    ids[rng.random(ids.shape) < 0.1] = -1
    sc = rng.standard_normal(ids.shape).astype(np.float32)
    sc[rng.random(ids.shape) < 0.2] = 0.5  # score ties between groups
    for prl, limit in ((60, 0), (0, 0), (400, 10), (1, 0)):
        ref = multivector_reduce(ids, sc, gid, prl=prl, limit=limit)
        res = st.multivector(torch.from_numpy(ids).cuda(), torch.from_numpy(sc).cuda(), prl=prl, limit=limit).cpu()
        _mv_compare(res, ref)
    st.close()


def test_search_multivector_end_to_end(cfg1, torch_cuda):
    """Config 3 in miniature: long queries x segments -> per-segment top-k -> union -> kbId groups,
    against the oracle chain exact_topk -> multivector_reduce."""
    torch = torch_cuda
    c, gid = cfg1["c"], cfg1["gid"]
    qs, _ = synth.make_queries(c, 6 * 4, seed=21)
    seg = qs.reshape(6, 4, 1536)
    res, sc, ids, flags = cfg1["store"].search_multivector(torch.from_numpy(seg).cuda(), 20, prl=60)
    torch.cuda.synchronize()
    ref_ids, ref_sc, _ = exact_topk_c(c, qs, 20)
    assert (ids.cpu().numpy().reshape(24, 20) == ref_ids).all()
    ref = multivector_reduce(ids.cpu().numpy(), sc.cpu().numpy(), gid, prl=60)
    _mv_compare(res.cpu(), ref)


def test_b200store_reference_store_kat(torch_cuda):
    """The reference's only numeric nearest-neighbour assertion
    (rag_engine/tests/test_storage_vector_store.py:10-24), against B200Store."""
    import asyncio

    from cmw_rag_b200 import B200Store

    async def run():
        store = B200Store(collection_name="test_collection", capacity=16)
        await store.add_async(texts=["a", "b"], metadatas=[{"kbId": "doc1"}, {"kbId": "doc2"}],
                              ids=["1", "2"], embeddings=[[0.1, 0.0, 0.0], [0.0, 0.1, 0.0]])
        results = await store.similarity_search_async(query_embedding=[0.1, 0.0, 0.0], k=1)
        assert len(results) == 1
        assert results[0].metadata["kbId"] == "doc1"
        # test_storage_vector_store.py:27-48: get / delete by where
        await store.add_async(texts=["c"], metadatas=[{"doc_stable_id": "d3", "kbId": "doc3"}], ids=["3"],
                              embeddings=[[0.0, 0.0, 0.1]])
        meta = await store.get_any_doc_meta_async({"doc_stable_id": "d3"})
        assert meta and meta["doc_stable_id"] == "d3"
        assert (await store.get_by_kb_id_async("doc2"))["kbId"] == "doc2"
        await store.delete_where_async({"doc_stable_id": "d3"})
        assert await store.get_any_doc_meta_async({"doc_stable_id": "d3"}) is None
        res = await store.similarity_search_async(query_embedding=[0.0, 0.0, 0.1], k=5)
        assert [r.metadata["kbId"] for r in res] == ["doc1", "doc2"]  # both score 0: lower id first
        # concurrent awaits coalesce into one launch (retriever.py:179-182 fan-out)
        before = store.stats["launch_batches"]
        outs = await asyncio.gather(*[store.similarity_search_async([0.1, 0.0, 0.0], k=2) for _ in range(4)])
        assert store.stats["launch_batches"] == before + 1 and store.stats["max_batch"] >= 4
        assert all(o[0].metadata["kbId"] == "doc1" for o in outs)
        col = await store.get_collection()
        raw = await col.query(query_embeddings=[[0.1, 0.0, 0.0]], n_results=2)
        assert raw["ids"] == [["1", "2"]] and abs(raw["distances"][0][0]) < 1e-6

    asyncio.run(run())


@pytest.mark.timeout(900)
def test_full_size_1m_properties(torch_cuda):
    """BASELINE.json config 2 size (1M x 1536, top-100): size-independent properties -- planted
    needles come back first, duplicated rows in id order, scores non-increasing, and a C-oracle
    cross-check on a few queries."""
    torch = torch_cuda
    from cmw_rag_b200 import DenseStore

    n, d, k = 1_000_000, 1536, 100
    g = torch.Generator(device="cuda").manual_seed(synth.CORPUS_SEED)
    st = DenseStore(d, n)
    block = 125_000
    host = np.empty((n, d), np.float32)
    for lo in range(0, n, block):
        x = torch.randn((block, d), generator=g, device="cuda", dtype=torch.float32)
        x = torch.nn.functional.normalize(x, dim=1)
        if lo == 0:
            keep17 = x[17].clone()
        if lo <= n // 2 < lo + block:
            x[n // 2 - lo] = keep17
        if lo + block == n:
            x[block - 1] = keep17
        st.append(x)
        host[lo:lo + block] = x.cpu().numpy()
    q, needle = synth.make_queries(host, 16)
    sc, ids, fl = st.search_host(q, k, mode="f32")
    planted = needle >= 0
    assert (ids[planted, 0] == needle[planted]).all()
    assert ids[1, :3].tolist() == [17, n // 2, n - 1]
    assert (np.diff(sc, axis=1) <= 0).all()
    assert (fl == 0).all()
    ref_ids, ref_sc, _ = exact_topk_c(host, q[:6], k)
    _check_exact(ids[:6], sc[:6], ref_ids, ref_sc)
    sc1, ids1, _ = st.search_host(q[:1], k, mode="f32", algo="scan")
    assert (ids1 == ids[:1]).all()
    scb, idsb, _ = st.search_host(q[:6], k, mode="bf16")
    recall = np.mean([len(set(idsb[b]) & set(ref_ids[b])) / k for b in range(6)])
    assert recall >= 0.9
    st.close()


@pytest.mark.parametrize("batch", [1, 7, 16, 33, 64, 128, 130, 192, 300])
def test_gemm_filter_exact(cfg1, batch):
    """K2 (tcgen05 GEMM filter) forced for every batch size: one partial group, exactly one group,
    several groups (300 -> 2 groups of 256 with padding)."""
    c = cfg1["c"]
    q, _ = synth.make_queries(c, batch, seed=31, tie_probe=False)
    ref_ids, ref_sc, _ = exact_topk_c(c, q, 20)
    sc, ids, fl = cfg1["store"].search_host(q, 20, mode="f32", algo="gemm")
    _check_exact(ids, sc, ref_ids, ref_sc)
    assert (fl == 0).all()
    sc, ids, fl = cfg1["store"].search_host(q, 20, mode="bf16", algo="gemm")
    recall = np.mean([len(set(ids[b]) & set(ref_ids[b])) / 20 for b in range(batch)])
    assert recall >= 0.97
    ex = np.einsum("bkd,bd->bk", c[ids].astype(np.float64), q.astype(np.float64))
    assert np.abs(sc - ex).max() <= BF16_TOL


def test_gemm_device_certificate_and_fallback(cfg1, torch_cuda):
    """An absurdly large certificate bound makes every certificate fail: the device API must flag the
    queries (never silently accept), the host API must repair them through the fp32 scan filter."""
    from cmw_rag_b200 import _native as N

    torch = torch_cuda
    q = cfg1["q"]
    old = N.get_option("bf16_eps")
    try:
        N.set_option("bf16_eps", 0.5)
        sc, ids, fl = cfg1["store"].search(torch.from_numpy(q).cuda(), 20, mode="f32", algo="gemm")
        torch.cuda.synchronize()
        assert int(fl.sum()) == q.shape[0]
        sc, ids, fl = cfg1["store"].search_host(q, 20, mode="f32", algo="gemm")
    finally:
        N.set_option("bf16_eps", old)
    _check_exact(ids, sc, cfg1["ref_ids"], cfg1["ref_sc"])
    assert (fl == 0).all()


def test_save_load_roundtrip_gpu(torch_cuda, tmp_path):
    """Persistence through the C ABI (cmw_store_read_rows_f32): a reloaded collection answers
    identically, tombstones included."""
    from cmw_rag_b200 import B200Store

    c = synth.make_corpus(3000, 96, seed=8)
    store = B200Store("persist", capacity=4096)
    store.add([f"t{i}" for i in range(3000)], [{"stable_id": f"{i}", "kbId": str(i // 8)} for i in range(3000)],
              ids=[str(i) for i in range(3000)], embeddings=c)
    store.delete(ids=["5", "17", "2999"])
    store.save(str(tmp_path / "col"))
    again = B200Store.load(str(tmp_path / "col"))
    q, _ = synth.make_queries(c, 9, seed=2)
    s1, i1, f1 = store.search(q, 30)
    s2, i2, f2 = again.search(q, 30)
    assert (i1 == i2).all() and np.array_equal(s1, s2) and again.count() == 2997
    live = np.ones(3000, bool)
    live[[5, 17, 2999]] = False
    ref_ids, ref_sc, _ = exact_topk_c(c, q, 30, live=live)
    _check_exact(i2, s2, ref_ids, ref_sc)


def test_clustered_corpus_exact_and_recall(torch_cuda):
    """Clustered corpus (4096-centroid style data has far denser score tails than iid
    Gaussian rows): exact mode must still return the oracle's ids (through the certificate or its
    fallback), bf16 mode is reported as recall."""
    torch = torch_cuda
    from cmw_rag_b200 import DenseStore

    n, d, k = 120_000, 256, 100
    c = synth.make_clustered_corpus(n, d, n_centroids=512, seed=77)
    q, _ = synth.make_queries(c, 48, seed=9, tie_probe=False)
    st = DenseStore(d, n)
    st.append(c)
    ref_ids, ref_sc, _ = exact_topk_c(c, q, k)
    sc, ids, fl = st.search_host(q, k, mode="f32")
    _check_exact(ids, sc, ref_ids, ref_sc)
    assert (fl == 0).all()
    _, _, fl_dev = st.search(torch.from_numpy(q).cuda(), k, mode="f32")
    torch.cuda.synchronize()
    # the device API may flag queries (dense tails); it must never return a wrong certified answer
    sc_d, ids_d, fl_d = st.search(torch.from_numpy(q).cuda(), k, mode="f32")
    torch.cuda.synchronize()
    good = fl_d.cpu().numpy() == 0
    assert (ids_d.cpu().numpy()[good] == ref_ids[good]).all()
    print("clustered: certified on device", int(good.sum()), "of", len(good))
    sc_b, ids_b, _ = st.search_host(q, k, mode="bf16")
    recall = np.mean([len(set(ids_b[b]) & set(ref_ids[b])) / k for b in range(q.shape[0])])
    assert recall >= 0.9, recall
    st.close()


def test_rerank_side_grouping_matches_reference_semantics(torch_cuda):
    """scored chunks -> articles (retriever.py:234-260,307-316) through K4 with one
    segment per query, against the CPU restatement; boosts as in reranker.py:165-181."""
    from cmw_rag_b200 import B200Store
    from cmw_rag_b200.articles import boost_and_order, group_scored_chunks
    from oracle.multivector import group_by_kbid, normalized_ranks

    rng = np.random.default_rng(12)
    n = 400
    kb = [("" if i % 37 == 0 else str(500 + i // 5) + ("-toc" if i % 11 == 0 else "")) for i in range(n)]
    store = B200Store("articles", capacity=512)
    store.add([f"c{i}" for i in range(n)],
              [{"stable_id": str(i), "kbId": kb[i], "has_code": i % 3 == 0, "tags": "x" if i % 4 == 0 else ""} for i in range(n)],
              ids=[str(i) for i in range(n)], embeddings=rng.standard_normal((n, 16)).astype(np.float32))
    rows = rng.permutation(n)[:60]
    raw = rng.random(60).astype(np.float32)
    weights = {"tag_match": 0.1, "code_presence": 0.05, "section_match": 0.0}
    order, final = boost_and_order(raw, [store._metas[r] for r in rows], weights, top_k=40)
    rows_o = rows[order]
    final32 = np.asarray(final, np.float32)
    for threshold in (None, 0.5):
        arts = group_scored_chunks(store, rows_o[None, :], final32[None, :], threshold=threshold)[0]
        gid = np.array([store.gid_for(k) for k in kb])
        ref = group_by_kbid(rows_o.tolist(), [float(x) for x in final32], gid)
        keep = [gi for gi in ref["order"] if threshold is None or ref["max"][gi] >= threshold]
        assert [a.kb_id for a in arts] == [store.key_of_gid(ref["gid"][gi]) for gi in keep]
        assert [a.score for a in arts] == [ref["max"][gi] for gi in keep]
        assert [a.rows for a in arts] == [[rows_o[m] for m in ref["members"][gi]] for gi in keep]
        assert [a.normalized_rank for a in arts] == normalized_ranks(len(keep))
        assert all(a.metadata["stable_id"] == str(a.rows[0]) for a in arts)


@pytest.mark.parametrize("k,batch", [(1, 3), (1, 40), (1000, 2), (257, 300)])
def test_extreme_k(cfg1, k, batch):
    """k = 1 and k near the K' ceiling (1024): the certificate has little or no head-room, so the host
    API may have to fall back -- the ids must be the oracle's either way."""
    c = cfg1["c"]
    q, _ = synth.make_queries(c, batch, seed=41 + k, tie_probe=False)
    ref_ids, ref_sc, _ = exact_topk_c(c, q, k)
    sc, ids, fl = cfg1["store"].search_host(q, k, mode="f32")
    _check_exact(ids, sc, ref_ids, ref_sc)
    assert (fl == 0).all()


def test_batch_larger_than_4096_and_store_growth(torch_cuda):
    from cmw_rag_b200 import B200Store

    c = synth.make_corpus(6000, 128, seed=3)
    store = B200Store("grow", capacity=1024)  # forces two growth steps
    for lo in range(0, 6000, 1500):
        store.add([f"t{i}" for i in range(lo, lo + 1500)], [{"kbId": str(i // 4)} for i in range(lo, lo + 1500)],
                  ids=[str(i) for i in range(lo, lo + 1500)], embeddings=c[lo:lo + 1500])
    assert store.count() == 6000 and store.dense.info()["capacity_rows"] >= 6000
    q, _ = synth.make_queries(c, 4200, seed=6, tie_probe=False)  # 4200 -> 17 query groups of 256
    ref_ids, ref_sc, _ = exact_topk_c(c, q, 10)
    sc, ids, fl = store.search(q, 10)
    _check_exact(ids, sc, ref_ids, ref_sc)


@pytest.mark.parametrize("strict", [1, 0])
def test_certificate_modes(cfg1, torch_cuda, strict):
    """strict_certificate = 1 (default): the rigorous residual bound behind the bf16 filter; 0: the statistical
    bound.  Same ids either way, every query certified by the DEVICE API (no repair chain) at this size."""
    from cmw_rag_b200 import _native as N

    torch = torch_cuda
    old = N.get_option("strict_certificate")
    assert old == 1.0, "the rigorous certificate must be the default"
    N.set_option("strict_certificate", strict)
    try:
        sc, ids, fl = cfg1["store"].search(torch.from_numpy(cfg1["q"]).cuda(), 20, mode="f32", algo="gemm")
        torch.cuda.synchronize()
    finally:
        N.set_option("strict_certificate", old)
    _check_exact(ids.cpu().numpy(), sc.cpu().numpy(), cfg1["ref_ids"], cfg1["ref_sc"])
    assert int(fl.sum()) == 0


def test_repair_chain_on_near_duplicates(torch_cuda):
    """600 near-duplicates of the best match (scores within 1e-6..1e-4 of each other): the bf16 filter
    cannot separate them, so the certificate behind K2 must fail on the device API -- and the host API's
    repair chain (stage 1: K' = 1024 on the overflow-proof schedule) must still return the oracle's ids."""
    torch = torch_cuda
    from cmw_rag_b200 import DenseStore

    n, d, k = 20000, 256, 20
    c = synth.make_corpus(n, d, seed=31, ties=False)
    rng = np.random.default_rng(2)
    base = c[7].copy()
    noise = rng.standard_normal((600, d)).astype(np.float32)
    c[1000:1600] = base[None, :] + 2e-4 * noise / np.sqrt(d)
    c[1000:1600] /= np.linalg.norm(c[1000:1600].astype(np.float64), axis=1, keepdims=True).astype(np.float32)
    q = np.stack([base + 0.05 * rng.standard_normal(d).astype(np.float32) / np.sqrt(d) for _ in range(8)])
    q2, _ = synth.make_queries(c, 8, seed=4, tie_probe=False)
    q = np.concatenate([q, q2]).astype(np.float32)
    st = DenseStore(d, n)
    st.append(c)
    ref_ids, ref_sc, _ = exact_topk_c(c, q, k)
    _, _, fl_dev = st.search(torch.from_numpy(q).cuda(), k, mode="f32", algo="gemm")
    torch.cuda.synchronize()
    assert int(fl_dev[:8].sum()) >= 1, "the near-duplicate queries should not be certifiable at K' = 96"
    sc, ids, fl = st.search_host(q, k, mode="f32", algo="gemm")
    _check_exact(ids, sc, ref_ids, ref_sc)
    assert (fl == 0).all()
    st.close()


def test_pipelined_host_api_matches_blocking(cfg1):
    """cmw_search_host_submit / _wait: several tickets in flight (pageable and page-locked buffers mixed,
    different batch sizes), waited for out of order, must return exactly what the blocking call returns."""
    from cmw_rag_b200 import _native as N
    from cmw_rag_b200.engine import pinned_empty

    st, q = cfg1["store"], cfg1["q"]
    parts = [q[:64], q[:1], q[5:38], q[10:26]]
    outs = [None, (pinned_empty((1, 20), np.float32), pinned_empty((1, 20), np.int64), np.zeros((1,), np.int32)), None, None]
    qin = [parts[0], parts[1], None, parts[3]]
    qin[2] = pinned_empty(parts[2].shape, np.float32)
    qin[2][:] = parts[2]
    tickets = [st.search_host_submit(qin[i], 20, mode="f32", out=outs[i]) for i in range(4)]
    assert sorted(tickets) == list(range(N.HOST_SLOTS))
    with pytest.raises(N.NativeError, match="slots are in flight"):
        st.search_host_submit(parts[0], 20, mode="f32")
    for i in (2, 0, 3, 1):
        sc, ids, fl = st.search_host_wait(tickets[i])
        lo = [0, 0, 5, 10][i]
        _check_exact(ids, sc, cfg1["ref_ids"][lo:lo + len(parts[i])], cfg1["ref_sc"][lo:lo + len(parts[i])])
        assert (fl == 0).all()
        if outs[i] is not None:
            assert ids is outs[i][1]
    with pytest.raises(N.NativeError, match="not in flight"):
        N.check(N.lib().cmw_search_host_wait(st._h, 0), "cmw_search_host_wait")
    # a steady pipeline: depth 2, 12 requests
    from collections import deque

    pending, done = deque(), 0
    for step in range(12):
        if len(pending) == 2:
            sc, ids, fl = st.search_host_wait(pending.popleft())
            _check_exact(ids, sc, cfg1["ref_ids"], cfg1["ref_sc"])
            done += 1
        pending.append(st.search_host_submit(q, 20, mode="f32"))
    while pending:
        sc, ids, fl = st.search_host_wait(pending.popleft())
        _check_exact(ids, sc, cfg1["ref_ids"], cfg1["ref_sc"])
        done += 1
    assert done == 12
    # and the blocking call still works next to it
    sc, ids, fl = st.search_host(q, 20, mode="f32")
    _check_exact(ids, sc, cfg1["ref_ids"], cfg1["ref_sc"])


def test_pipelined_host_api_from_several_threads(cfg1):
    """Four host threads, each keeping one ticket in flight on the same store (the header's threading
    contract for the pipelined entry points), next to a fifth thread issuing blocking calls."""
    import threading

    st, q = cfg1["store"], cfg1["q"]
    errors = []

    def worker(tid):
        try:
            lo = 8 * tid
            for it in range(25):
                if tid < 4:
                    t = st.search_host_submit(q[lo:lo + 8 + it % 3], 20, mode="f32")
                    sc, ids, fl = st.search_host_wait(t)
                else:
                    sc, ids, fl = st.search_host(q[lo:lo + 8 + it % 3], 20, mode="f32")
                n = 8 + it % 3
                _check_exact(ids, sc, cfg1["ref_ids"][lo:lo + n], cfg1["ref_sc"][lo:lo + n])
                assert (fl == 0).all()
        except Exception as exc:  # noqa: BLE001 - reported below
            errors.append((tid, repr(exc)))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(5)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


def test_pipelined_host_api_repairs_flagged_queries(torch_cuda):
    """A ticket whose queries fail the certificate goes through the repair chain inside the wait."""
    from cmw_rag_b200 import DenseStore

    n, d, k = 20000, 256, 20
    c = synth.make_corpus(n, d, seed=31, ties=False)
    rng = np.random.default_rng(2)
    base = c[7].copy()
    noise = rng.standard_normal((600, d)).astype(np.float32)
    c[1000:1600] = base[None, :] + 2e-4 * noise / np.sqrt(d)
    c[1000:1600] /= np.linalg.norm(c[1000:1600].astype(np.float64), axis=1, keepdims=True).astype(np.float32)
    q = np.stack([base + 0.05 * rng.standard_normal(d).astype(np.float32) / np.sqrt(d) for _ in range(8)])
    q2, _ = synth.make_queries(c, 8, seed=4, tie_probe=False)
    q = np.concatenate([q, q2]).astype(np.float32)
    st = DenseStore(d, n)
    st.append(c)
    ref_ids, ref_sc, _ = exact_topk_c(c, q, k)
    t0 = st.search_host_submit(q, k, mode="f32", algo="gemm")
    t1 = st.search_host_submit(q[8:], k, mode="f32", algo="gemm")
    sc, ids, fl = st.search_host_wait(t0)
    _check_exact(ids, sc, ref_ids, ref_sc)
    assert (fl == 0).all()
    sc, ids, fl = st.search_host_wait(t1)
    _check_exact(ids, sc, ref_ids[8:], ref_sc[8:])
    st.close()


def _fuzz_case(seed: int):
    """One random configuration: shape, metric, duplicates, zero rows, tombstones, several appends."""
    rng = np.random.default_rng(1000 + seed)
    d = int(rng.choice([8, 16, 24, 64, 72, 128, 200, 256, 384, 512, 768, 1000, 1024, 1536, 2048]))
    n = int(rng.choice([1, 2, 31, 127, 128, 129, 255, 257, 1000, 4095, 4096, 4097, 8191, 12289, 30011]))
    if n * d > 24_000_000:
        n = 24_000_000 // d
    b = int(rng.choice([1, 2, 3, 15, 16, 17, 31, 33, 64, 100, 255, 256, 257, 300, 513]))
    k = int(rng.choice([1, 2, 5, 20, 50, 100, 130, 257, 600]))
    metric = "ip" if rng.random() < 0.3 else "cosine"
    c = rng.standard_normal((n, d)).astype(np.float32)
    if metric == "ip" or rng.random() < 0.5:
        c *= rng.uniform(0.2, 3.0, size=(n, 1)).astype(np.float32)  # un-normalised rows
    else:
        c /= np.linalg.norm(c, axis=1, keepdims=True)
    if n >= 8:  # exact duplicates (ties -> lower id first) and a zero row
        for _ in range(int(rng.integers(1, 6))):
            src, dst = rng.integers(0, n, size=2)
            c[dst] = c[src]
        c[int(rng.integers(0, n))] = 0
    q = rng.standard_normal((b, d)).astype(np.float32)
    if n >= 8:
        pick = rng.integers(0, n, size=b)
        planted = rng.random(b) < 0.5
        q[planted] = c[pick[planted]] + 0.3 * q[planted] / np.sqrt(d)
    live = np.ones(n, bool)
    if n >= 16 and rng.random() < 0.6:
        live[rng.integers(0, n, size=max(1, n // 10))] = False
    splits = sorted(set(rng.integers(0, n + 1, size=int(rng.integers(0, 3))).tolist()) | {0, n})
    return dict(d=d, n=n, b=b, k=k, metric=metric, c=c, q=q, live=live, splits=splits)


@pytest.mark.parametrize("seed", range(36))
def test_fuzz_search_against_oracle(torch_cuda, seed):
    """Random shapes / metrics / duplicates / tombstones through the host C ABI (default algorithm choice and
    both explicit filters): ids identical to the fp64 oracle, scores within 1e-5 relative to the score scale."""
    from cmw_rag_b200 import DenseStore

    f = _fuzz_case(seed)
    c, q, live, n, k = f["c"], f["q"], f["live"], f["n"], f["k"]
    st = DenseStore(f["d"], n + int(seed % 3))
    for lo, hi in zip(f["splits"][:-1], f["splits"][1:]):
        st.append(c[lo:hi])
    dead = np.flatnonzero(~live)
    if dead.size:
        st.tombstone(dead)
    assert st.live_rows == int(live.sum())
    ref_ids, ref_sc, _ = exact_topk_c(c, q, k, metric=f["metric"], live=live)
    scale = max(1.0, float(np.abs(ref_sc[np.isfinite(ref_sc)]).max())) if np.isfinite(ref_sc).any() else 1.0
    for algo in ("auto", "scan", "gemm"):
        sc, ids, fl = st.search_host(q, k, metric=f["metric"], mode="f32", algo=algo)
        # a tie group larger than the candidate set may stay flagged (known limit); never here
        assert (fl == 0).all(), (algo, f["d"], n, f["b"], k)
        assert (ids == ref_ids).all(), (algo, f["d"], n, f["b"], k, f["metric"], int((ids != ref_ids).sum()))
        fin = np.isfinite(ref_sc)
        assert (np.isneginf(sc) == np.isneginf(ref_sc)).all()
        if fin.any():
            assert np.abs(sc[fin] - ref_sc[fin]).max() <= F32_TOL * scale
    # bf16 mode: same candidate set up to rounding -- every returned id is live and in range, scores close
    sc, ids, fl = st.search_host(q, k, metric=f["metric"], mode="bf16")
    ok = ids >= 0
    assert live[ids[ok]].all()
    for row in ids:
        got = row[row >= 0]
        assert np.unique(got).size == got.size, "duplicate ids in a result row"
    assert ok.sum(axis=1).min() == min(k, int(live.sum()))
    st.close()


@pytest.mark.parametrize("b", [257, 320, 384, 400, 513, 600, 777, 1000])
def test_query_groups_only_as_wide_as_needed(torch_cuda, b):
    """Batches above 256 are split into ceil(B / 256) query groups of equal width in steps of 64 (384 -> 2 x 192, not
    2 x 256 with a quarter of the tensor work on padding): CTA-pair and 1-CTA kernels, fp16 tiles and the tf32
    filter of a store without 16-bit tiles, all against the oracle."""
    from cmw_rag_b200 import DenseStore
    from cmw_rag_b200 import _native as N

    n, d, k = 20000, 256, 30
    c = synth.make_corpus(n, d, seed=31)
    q, _ = synth.make_queries(c, b, seed=32)
    ref_ids, ref_sc, _ = exact_topk_c(c, q, k)
    st = DenseStore(d, n)
    st.append(c)
    st32 = DenseStore(d, n, f32=True, bf16=False)
    st32.append(c)
    try:
        for pairs in (1, 0):
            N.set_option("gemm_2cta", pairs)
            sc, ids, fl = st.search_host(q, k, mode="f32", algo="gemm")
            assert (fl == 0).all() and (ids == ref_ids).all(), (b, pairs, int((ids != ref_ids).sum()))
            assert np.abs(sc - ref_sc).max() <= F32_TOL
        sc, ids, fl = st32.search_host(q, k, mode="f32")
        assert (fl == 0).all() and (ids == ref_ids).all(), (b, "tf32")
    finally:
        N.set_option("gemm_2cta", 1)
        st.close()
        st32.close()


@pytest.mark.parametrize("seed", range(16))
def test_fuzz_two_phase_row_shards_on_one_gpu(torch_cuda, seed):
    """The row-sharded search (filter | k-th over all shards | finish with the global cut | merge + cross-shard
    certificate) over G shard stores on ONE GPU, on the fuzz cases above: random shapes and metrics, exact
    duplicates that land in DIFFERENT shards (the lower global id must win across shards), tombstones, shards that
    are short, hold fewer than k rows, or are empty (more ranks than rows).  Unflagged answers must be the fp64
    oracle's; with so few rows per shard nothing may stay flagged here."""
    import torch

    from cmw_rag_b200 import DenseStore
    from cmw_rag_b200.engine import shard_kth, shard_merge
    from cmw_rag_b200.sharded import shard_bounds

    f = _fuzz_case(100 + seed)
    c, q, live, n, d = f["c"], f["q"], f["live"], f["n"], f["d"]
    k = min(f["k"], 257)
    q = q[:64]
    b = q.shape[0]
    G = int(np.random.default_rng(seed).choice([1, 2, 3, 5, 8]))
    if G * k > 8192:
        G = max(1, 8192 // k)
    bounds = shard_bounds(n, G)
    shards = []
    for lo, hi in bounds:
        st = DenseStore(d, max(1, hi - lo), id_offset=lo)
        if hi > lo:
            st.append(c[lo:hi])
            dead = np.flatnonzero(~live[lo:hi])
            if dead.size:
                st.tombstone(dead)
        shards.append(st)
    ref_ids, ref_sc, _ = exact_topk_c(c, q, k, metric=f["metric"], live=live)
    scale = max(1.0, float(np.abs(ref_sc[np.isfinite(ref_sc)]).max())) if np.isfinite(ref_sc).any() else 1.0
    qd = torch.from_numpy(q).cuda()
    for algo in ("auto", "scan"):
        kw = dict(metric=f["metric"], mode="f32", algo=algo)
        ftop = torch.stack([st.search_filter(qd, k, **kw) for st in shards])
        kth = shard_kth(ftop, k)
        blocks = torch.cat([st.search_finish(qd, k, global_kth=kth, **kw) for st in shards])
        ms, mi, _, fl = shard_merge(blocks, G, b, k)
        torch.cuda.synchronize()
        ids, sc, fl = mi.cpu().numpy(), ms.cpu().numpy(), fl.cpu().numpy()
        assert (fl == 0).all(), (algo, G, n, d, b, k, f["metric"], int((fl != 0).sum()))
        assert (ids == ref_ids).all(), (algo, G, n, d, b, k, f["metric"], int((ids != ref_ids).sum()))
        fin = np.isfinite(ref_sc)
        assert (np.isneginf(sc) == np.isneginf(ref_sc)).all()
        if fin.any():
            assert np.abs(sc[fin] - ref_sc[fin]).max() <= F32_TOL * scale
    for st in shards:
        st.close()


def test_b200store_compaction_gpu(torch_cuda):
    """B200Store.compact(): tombstoned rows are physically dropped; results (as stable ids and scores) are
    those of the oracle over the live rows before and after."""
    from cmw_rag_b200 import B200Store

    n, d, k = 6000, 64, 10
    c = synth.make_corpus(n, d, seed=77, ties=False)
    q, _ = synth.make_queries(c, 9, seed=78, tie_probe=False)
    store = B200Store("cmp", capacity=8192)
    store.add([f"t{i}" for i in range(n)], [{"doc_stable_id": f"D{i // 3}", "kbId": str(i // 3)} for i in range(n)],
              ids=[f"c{i}" for i in range(n)], embeddings=c)
    for doc in range(0, 2000, 2):
        store.delete(where={"doc_stable_id": f"D{doc}"})
    live = np.ones(n, bool)
    for doc in range(0, 2000, 2):
        live[3 * doc:3 * doc + 3] = False
    assert store.count() == int(live.sum()) == 3000
    ref_ids, ref_sc, _ = exact_topk_c(c, q, k, live=live)
    want = [[f"c{r}" for r in row] for row in ref_ids]
    res = store.query(q, k)
    assert res["ids"] == want
    assert store.compact() == 3000 and store.dense.rows == 3000
    res2 = store.query(q, k)
    assert res2["ids"] == want
    assert np.allclose(np.array(res2["distances"]), 1.0 - ref_sc, atol=F32_TOL)
    assert np.array_equal(np.array(res2["distances"]), np.array(res["distances"]))
    store.add(["z"], [{"doc_stable_id": "Z", "kbId": "77"}], ids=["z"], embeddings=q[:1])
    assert store.query(q[:1], 1)["ids"] == [["z"]]


def test_repeated_host_searches_track_store_changes(torch_cuda):
    """The same small blocking host search repeated across tombstones, appends (new slab schedule) and option
    changes keeps returning the oracle's answer for the store as it is at that moment."""
    from cmw_rag_b200 import DenseStore
    from cmw_rag_b200 import _native as N

    n, d, k = 9000, 128, 10
    c = synth.make_corpus(n, d, seed=91, ties=False)
    q, _ = synth.make_queries(c, 3, seed=92, tie_probe=False)
    st = DenseStore(d, n + 500)
    st.append(c[:8000])
    live = np.ones(8000, bool)
    ref_ids, ref_sc, _ = exact_topk_c(c[:8000], q, k)
    for call in range(3):
        sc, ids, fl = st.search_host(q, k)
        _check_exact(ids, sc, ref_ids, ref_sc)
        sc1, ids1, _ = st.search_host(q[:1], k)
        _check_exact(ids1, sc1, ref_ids[:1], ref_sc[:1])
    dead = ref_ids[:, 0].tolist()
    st.tombstone(dead)
    live[dead] = False
    ref2, ref2_sc, _ = exact_topk_c(c[:8000], q, k, live=live)
    sc, ids, fl = st.search_host(q, k)
    _check_exact(ids, sc, ref2, ref2_sc)
    st.append(c[8000:])
    live = np.concatenate([live, np.ones(n - 8000, bool)])
    ref3, ref3_sc, _ = exact_topk_c(c, q, k, live=live)
    for algo in ("auto", "scan", "gemm"):
        for call in range(2):
            sc, ids, fl = st.search_host(q, k, algo=algo)
            _check_exact(ids, sc, ref3, ref3_sc)
    N.set_option("gemm_clc", 0)
    try:
        sc, ids, fl = st.search_host(np.repeat(q, 100, axis=0), k)
        _check_exact(ids, sc, np.repeat(ref3, 100, axis=0), np.repeat(ref3_sc, 100, axis=0))
    finally:
        N.set_option("gemm_clc", 1)
    N.set_option("scan_permute", 0)
    try:
        for qq in (q, np.repeat(q, 100, axis=0)):
            sc, ids, fl = st.search_host(qq, k)
            reps = qq.shape[0] // 3
            _check_exact(ids, sc, np.repeat(ref3, reps, axis=0), np.repeat(ref3_sc, reps, axis=0))
    finally:
        N.set_option("scan_permute", 1)
    st.close()


def test_pure_c_client(tmp_path):
    """The C-ABI boundary used from plain C (examples/c_client.c): no CUDA headers, no Python objects."""
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib_dir = os.path.join(root, "cmw_rag_b200", "csrc")
    exe = str(tmp_path / "c_client")
    subprocess.run(["gcc", "-O2", "-I" + os.path.join(root, "include"), os.path.join(root, "examples", "c_client.c"),
                    "-o", exe, "-L" + lib_dir, "-lcmwdense", "-Wl,-rpath," + lib_dir, "-lm"], check=True)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "c_client ok" in res.stdout


# ------------------------------------------------------------------------------------------------
# round 2: tombstoned slabs, the rigorous certificate, K1 width
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,dead", [(6000, 4096), (7000, 5000), (20000, 4096), (20000, 9000), (100000, 65536),
                                    (100000, 70000)])
def test_leading_rows_all_tombstoned(torch_cuda, n, dead):
    """A re-indexed collection: the first `dead` rows are tombstones, the live rows were appended behind them.
    The dense first slab (4096 rows, or 65536 for small batches) then sees no live row at all; its -inf
    placeholders must not count as candidates (they used to fill the pools and push every later admission off
    the end).  Device API, no repair chain: ids identical to the oracle, nothing flagged."""
    torch = torch_cuda
    from cmw_rag_b200 import DenseStore
    from cmw_rag_b200 import _native as N

    d, k = 128, 10
    c = synth.make_corpus(n, d, seed=100 + n % 97, ties=False)
    q, _ = synth.make_queries(c[dead:], 40, seed=3, tie_probe=False)
    live = np.ones(n, np.uint8)
    live[:dead] = 0
    st = DenseStore(d, n)
    st.append(c)
    st.tombstone(np.arange(dead))
    ref_ids, ref_sc, _ = exact_topk_c(c, q, k, live=live)
    qd = torch.from_numpy(q).cuda()
    old = N.get_option("wide_dense")
    try:
        for wide in (1, 0):
            N.set_option("wide_dense", wide)
            for algo in ("scan", "gemm"):
                for b in (3, 40):
                    sc, ids, fl = st.search(qd[:b], k, mode="f32", algo=algo)
                    torch.cuda.synchronize()
                    _check_exact(ids.cpu().numpy(), sc.cpu().numpy(), ref_ids[:b], ref_sc[:b])
                    assert int(fl.sum()) == 0, (wide, algo, b)
    finally:
        N.set_option("wide_dense", old)
    st.close()


def test_scattered_tombstones_half_dead(torch_cuda):
    """Every second row dead, plus one dead block in the middle: the live-row estimate of the slab schedule
    (store-wide live fraction for the permuted K2 scan, exact per-block counts for K1) must keep the pools
    from overflowing."""
    torch = torch_cuda
    from cmw_rag_b200 import DenseStore

    n, d, k = 60000, 128, 20
    c = synth.make_corpus(n, d, seed=77, ties=False)
    live = np.ones(n, np.uint8)
    live[::2] = 0
    live[20000:30000] = 0
    q, _ = synth.make_queries(c[live.astype(bool)], 48, seed=5, tie_probe=False)
    st = DenseStore(d, n)
    st.append(c)
    st.tombstone(np.flatnonzero(live == 0))
    ref_ids, ref_sc, _ = exact_topk_c(c, q, k, live=live)
    qd = torch.from_numpy(q).cuda()
    for algo in ("scan", "gemm"):
        for b in (2, 48):
            sc, ids, fl = st.search(qd[:b], k, mode="f32", algo=algo)
            torch.cuda.synchronize()
            _check_exact(ids.cpu().numpy(), sc.cpu().numpy(), ref_ids[:b], ref_sc[:b])
            assert int(fl.sum()) == 0, (algo, b)
    st.close()


def _bf16_round(x32: np.ndarray) -> np.ndarray:
    """fp32 -> bf16 (round to nearest even) -> fp32, the way __floats2bfloat162_rn does."""
    bits = np.ascontiguousarray(x32, dtype=np.float32).view(np.uint32).astype(np.uint64)
    bits = bits + 0x7FFF + ((bits >> 16) & 1)
    return (bits & 0xFFFF0000).astype(np.uint32).view(np.float32)


def _structured_case(d, m, n_decoys, seed):
    """Rows and queries whose bf16 roundings all go the same way (m equal-magnitude non-zeros: 1/sqrt(1017) =
    1.00344 * 2^-5 loses 0.34 % in every element, 1/sqrt(254) gains 0.39 %) next to dense rows that round
    randomly.  The structured rows' filter scores are therefore shifted by up to 7e-3 against the dense ones.
    A few dense "decoys" sit within that shift of the structured rows' exact scores -- sparse enough that the
    K'-th candidate is far below -- so a certificate that assumes independent rounding errors (round 1's)
    accepts a top-k from which a structured row is missing, or in which one is wrongly present."""
    rng = np.random.default_rng(seed)
    support = rng.permutation(d)[:m]
    base = np.zeros(d, np.float64)
    base[support] = 1.0 / np.sqrt(m)
    f0 = max(2, m // 50)
    flips = np.arange(f0, f0 + 8)
    fam = np.tile(base, (flips.size, 1))
    for i, f in enumerate(flips):
        fam[i, support[rng.permutation(m)[:f]]] *= -1.0
    cos_band = 1.0 - 2.0 * flips / m
    # sparse decoys either side of the structured rows' exact scores: cos(theta) * base + sin(theta) * noise
    ct = np.concatenate([np.linspace(c * (1 - 0.0045), c * (1 + 0.0045), 7) for c in cos_band[3:]])
    ct = np.unique(np.round(ct[ct < 0.9999], 6))
    noise = rng.standard_normal((ct.size, d))
    noise -= (noise @ base)[:, None] * base[None, :]
    noise /= np.linalg.norm(noise, axis=1, keepdims=True)
    decoys = ct[:, None] * base[None, :] + np.sqrt(1.0 - ct ** 2)[:, None] * noise
    filler = rng.standard_normal((n_decoys, d))
    filler /= np.linalg.norm(filler, axis=1, keepdims=True)
    rows = np.concatenate([filler[: n_decoys // 2], decoys, fam, filler[n_decoys // 2:]])
    rows = rows[rng.permutation(rows.shape[0])].astype(np.float32)
    # queries: the structured direction itself, and variants with a few flipped signs (still structured)
    qs = [base.copy()]
    for f in (1, 2, 3):
        v = base.copy()
        v[support[rng.permutation(m)[:f]]] *= -1.0
        qs.append(v)
    return rows, np.asarray(qs, np.float32)


@pytest.mark.parametrize("d,m", [(1536, 1017), (1536, 254), (4096, 4068), (256, 254)])
@pytest.mark.parametrize("metric", ["cosine", "ip"])
@pytest.mark.parametrize("tiles16", ["bf16", "f16", "tf32"])
def test_certificate_is_sound_on_adversarial_roundings(torch_cuda, d, m, metric, tiles16):
    """The guarantee of CMW_MODE_F32_EXACT: a query comes back with the oracle's ids OR flagged -- never a
    silent miss -- whatever the rounding structure of the vectors.  (Round 1's bound assumed u = 2^-9 and
    independent roundings; on these inputs the bf16 filter is off by up to 6.6e-3 and it certified wrong
    answers.)  The host API must then repair the flagged queries to the oracle's answer."""
    torch = torch_cuda
    from cmw_rag_b200 import DenseStore
    from cmw_rag_b200 import _native as N

    rows, q = _structured_case(d, m, n_decoys=3000, seed=d + m)
    if metric == "ip":
        rows = rows * np.linspace(0.97, 1.03, rows.shape[0], dtype=np.float32)[:, None]
        q = q * np.float32(1.7)
    # "tf32": no 16-bit tiles at all -- the filter reads the fp32 rows through kind::tf32 MMAs, which drop 13
    # mantissa bits of every element the same way (truncation): the same kind of systematic error
    st = DenseStore(d, rows.shape[0], tiles16=tiles16) if tiles16 != "tf32" else DenseStore(d, rows.shape[0], bf16=False)
    st.append(rows)
    qd = torch.from_numpy(q).cuda()
    silent, flagged = 0, 0
    for k in (6, 10, 14, 20):
        ref_ids, ref_sc, _ = exact_topk_c(rows, q, k, metric=metric)
        sc, ids, fl = st.search(qd, k, metric=metric, mode="f32", algo="gemm")
        torch.cuda.synchronize()
        ids, fl = ids.cpu().numpy(), fl.cpu().numpy()
        for b in range(q.shape[0]):
            same = (ids[b] == ref_ids[b]).all()
            flagged += int(fl[b] != 0)
            if not same and fl[b] == 0:
                silent += 1
        sc_h, ids_h, fl_h = st.search_host(q, k, metric=metric, mode="f32", algo="gemm")
        assert (ids_h == ref_ids).all(), (k, "host API (repair chain) must return the oracle's ids")
        assert np.abs(sc_h - ref_sc).max() <= F32_TOL * (3.0 if metric == "ip" else 1.0)
        assert (fl_h == 0).all()
    assert silent == 0, f"{silent} queries returned wrong ids without CMW_FLAG_UNCERTIFIED"
    # the statistical bound is documented as breakable by exactly this: make sure the test would have caught it
    old = N.get_option("strict_certificate")
    N.set_option("strict_certificate", 0)
    try:
        stat_silent = 0
        for k in (6, 10, 14, 20):
            ref_ids, _, _ = exact_topk_c(rows, q, k, metric=metric)
            _, ids, fl = st.search(qd, k, metric=metric, mode="f32", algo="gemm")
            torch.cuda.synchronize()
            stat_silent += int(sum((ids.cpu().numpy()[b] != ref_ids[b]).any() and int(fl[b]) == 0
                                   for b in range(q.shape[0])))
    finally:
        N.set_option("strict_certificate", old)
    print(f"adversarial {tiles16} d={d} m={m} {metric}: rigorous flagged {flagged}, "
          f"statistical silent misses {stat_silent}")
    st.close()


def _f16_round(x32: np.ndarray) -> np.ndarray:
    """fp32 -> fp16 (round to nearest even), results below the smallest normal flushed to zero -> fp32."""
    h = np.ascontiguousarray(x32, dtype=np.float32).astype(np.float16).astype(np.float32)
    h[np.abs(h) < 2.0 ** -14] = 0.0
    return h


@pytest.mark.parametrize("tiles16", ["bf16", "f16"])
def test_filter_error_stays_inside_the_rigorous_bound(torch_cuda, tiles16):
    """What the certificate rests on, measured: bf16-mode scores (the tensor-core filter's own output) against
    (a) the fp64 dot product of the bf16-ROUNDED operands -- the fp32 accumulation error, bounded by D * 2^-23 --
    and (b) the exact cosine -- bounded by r_q + (1 + r_q) R_c + D * 2^-23 with the residuals recomputed here
    from the same rounding rule."""
    torch = torch_cuda
    from cmw_rag_b200 import DenseStore

    d, k = 1536, 64
    rows_a, q_a = _structured_case(d, 1017, n_decoys=1500, seed=9)
    c = np.concatenate([rows_a, synth.make_corpus(20000, d, seed=8, ties=False)])
    q = np.concatenate([q_a, synth.make_queries(c, 28, seed=9, tie_probe=False)[0]])
    st = DenseStore(d, c.shape[0], tiles16=tiles16)
    st.append(c)
    rnd = _bf16_round if tiles16 == "bf16" else _f16_round
    u = 2.0 ** -8 if tiles16 == "bf16" else 2.0 ** -11
    sc, ids, _ = st.search(torch.from_numpy(q).cuda(), k, mode="bf16", algo="gemm")
    torch.cuda.synchronize()
    sc, ids = sc.cpu().numpy().astype(np.float64), ids.cpu().numpy()
    c64, q64 = c.astype(np.float64), q.astype(np.float64)
    c_hat = c64 / np.linalg.norm(c64, axis=1, keepdims=True)
    q_hat = q64 / np.linalg.norm(q64, axis=1, keepdims=True)
    c_t = rnd(c_hat.astype(np.float32)).astype(np.float64)
    q_t = rnd(q_hat.astype(np.float32)).astype(np.float64)
    r_c = np.linalg.norm(c_hat - c_t, axis=1)
    r_q = np.linalg.norm(q_hat - q_t, axis=1)
    # round to nearest: relative error <= u per element (+ the flushed fp16 elements, each below 2^-14)
    assert r_c.max() <= u + 2.0 ** -14 * np.sqrt(d) and r_q.max() <= u + 2.0 ** -14 * np.sqrt(d)
    slack = d * 2.0 ** -23
    worst_acc, worst_tot = 0.0, 0.0
    for b in range(q.shape[0]):
        rounded = c_t[ids[b]] @ q_t[b]
        exact = c_hat[ids[b]] @ q_hat[b]
        worst_acc = max(worst_acc, float(np.abs(sc[b] - rounded).max()))
        bound = r_q[b] + (1 + r_q[b]) * r_c.max() + slack
        assert np.abs(sc[b] - exact).max() <= bound, (b, np.abs(sc[b] - exact).max(), bound)
        worst_tot = max(worst_tot, float(np.abs(sc[b] - exact).max() / bound))
    assert worst_acc <= slack, (worst_acc, slack)
    print(f"{tiles16}: tensor-pipe accumulation error max {worst_acc:.3e} (slack {slack:.3e}); "
          f"filter error / rigorous bound max {worst_tot:.3f}; residuals r_c max {r_c.max():.3e} "
          f"r_q max {r_q.max():.3e}")
    st.close()


@pytest.mark.parametrize("batch", [3, 4, 5, 8, 9, 16])
def test_scan_kernel_four_queries_per_pass(cfg1, batch):
    """K1 takes up to 4 queries per pass over the corpus (3-4: queries in shared memory)."""
    sc, ids, fl = cfg1["store"].search_host(cfg1["q"][:batch], 20, mode="f32", algo="scan")
    _check_exact(ids, sc, cfg1["ref_ids"][:batch], cfg1["ref_sc"][:batch])
    assert (fl == 0).all()


@pytest.mark.parametrize("metric", ["cosine", "ip"])
def test_bf16_tiles_store(cfg1, torch_cuda, metric):
    """BASELINE.json's literal tile format: a store whose 16-bit tiles are bf16 (the default is fp16).  Exact mode
    through the tensor-core filter (batches on the 1-CTA and the CTA-pair kernel) and the 16-bit scan, device API:
    the oracle's ids, nothing flagged; approximate mode within the north star's 2e-3.  Also a badly scaled query
    (norm 1e-6 / 1e+6): the filter normalises the query in both metrics, so neither tile format cares."""
    torch = torch_cuda
    from cmw_rag_b200 import DenseStore

    c, q = cfg1["c"], cfg1["q"].copy()
    q[3] *= 1e-6
    q[4] *= 1e6
    st = DenseStore(1536, c.shape[0], tiles16="bf16")
    st.append(c)
    assert st.tiles16 == "bf16"
    k = 20
    ref_ids, ref_sc, _ = exact_topk_c(c, q, k, metric=metric)
    qd = torch.from_numpy(q).cuda()
    scale = np.maximum(1.0, np.linalg.norm(q.astype(np.float64), axis=1))[:, None] if metric == "ip" else 1.0
    for b in (1, 40, 64):
        sc, ids, fl = st.search(qd[:b], k, metric=metric, mode="f32", algo="gemm")
        torch.cuda.synchronize()
        assert (ids.cpu().numpy() == ref_ids[:b]).all() and int(fl.sum()) == 0, b
        err = np.abs(sc.cpu().numpy() - ref_sc[:b]) / (scale[:b] if metric == "ip" else 1.0)
        assert err.max() <= F32_TOL
    for store in (st, cfg1["store"]):  # bf16 tiles, fp16 tiles
        for algo in ("scan", "gemm"):
            sc, ids, _ = store.search_host(q, k, metric=metric, mode="bf16", algo=algo)
            recall = np.mean([len(set(ids[b]) & set(ref_ids[b])) / k for b in range(q.shape[0])])
            assert recall >= 0.97, (algo, recall)
            ex = np.einsum("bkd,bd->bk", c[ids].astype(np.float64), q.astype(np.float64))
            if metric == "cosine":
                ex /= np.linalg.norm(q.astype(np.float64), axis=1)[:, None]
                ex /= np.linalg.norm(c[ids].astype(np.float64), axis=2)
            assert (np.abs(sc - ex) / scale).max() <= BF16_TOL, algo
    st.close()


# ------------------------------------------------------------------------------------------------
# round 2: rerank-side plumbing, golden replay through the CUDA backend, batching front-end on the GPU
# ------------------------------------------------------------------------------------------------
def test_tool_result_merge_matches_reference_golden(golden_dir, torch_cuda):
    """accumulate_articles_from_tool_results (rag_engine/tools/utils.py:70-152) -- dedup by kb_id keeping the best
    score, sort best first -- through K4, against lists the reference's own function produced."""
    from cmw_rag_b200.articles import merge_tool_results

    with open(os.path.join(golden_dir, "f3_golden.json")) as f:
        g = json.load(f)
    assert len(g["tool_merge"]) >= 20
    for case in g["tool_merge"]:
        lists = []
        for raw in case["tool_results"]:
            arts = json.loads(raw)["articles"]
            lists.append([(a["kb_id"], a["metadata"].get("rerank_score"), a["content"]) for a in arts])
        got = merge_tool_results(lists)
        assert [[kb, content, score] for kb, score, content in got] == case["expected"]
    assert merge_tool_results([]) == [] and merge_tool_results([[], []]) == []


def _golden_store(golden, torch):
    from cmw_rag_b200.store import B200Store

    n, d = golden["n"], 48
    corpus = synth.make_corpus(n, d, seed=99)
    kb = golden["kb"]
    store = B200Store(collection_name="golden", capacity=n)
    store.add(texts=[f"chunk {r}" for r in range(n)],
              metadatas=[{"stable_id": f"{r:012d}", "kbId": kb[r]} for r in range(n)],
              ids=[f"{r:012d}" for r in range(n)], embeddings=corpus)
    return store


def test_golden_replay_through_the_cuda_backend(golden_dir, torch_cuda):
    """The drop-in check that needs no reference checkout: the embeddings the reference's unmodified RAGRetriever
    handed to its store and the Article lists it returned were recorded by tests/golden/make_golden.py; here the
    same embeddings go through B200Store on the GPU (batched search, K4 union / cap, the reference's rerank stand-in
    or its no-rerank truncation, K4 grouping, inclusive threshold, ranks) and must reproduce those lists."""
    torch = torch_cuda
    from cmw_rag_b200.articles import boost_and_order, group_scored_chunks

    with open(os.path.join(golden_dir, "multivector_golden.json")) as f:
        golden = json.load(f)
    store = _golden_store(golden, torch)
    assert len(golden["cases"]) == 7
    for case in golden["cases"]:
        p = case["params"]
        segs = case["segments"]
        assert all("query" in s for s in segs), "golden fixture predates the recorded embeddings"
        qv = np.asarray([s["query"] for s in segs], np.float32)[None, :, :]  # [1, S, d]
        k = p["top_k_retrieve"]
        res, ids, scores = store.search_multivector(qv, k, prl=p["prl"])
        # (1) per-segment lists = what the reference's store returned
        for si, s_ in enumerate(segs):
            assert ids[0, si].tolist() == s_["ids"], (case["name"], si)
            assert np.abs(scores[0, si] - np.asarray(s_["scores"], np.float32)).max() <= F32_TOL
        cn = int(res.cand_n[0])
        cand = res.cand_ids[0, :cn].numpy()
        cand_sc = res.cand_scores[0, :cn].numpy()
        if p["rerank"]:
            # (2) the union handed to the reranker, in the reference's first-seen order
            assert [f"{int(r):012d}" for r in cand] == case["rerank_input_stable_ids"], case["name"]
            # the fixture's stand-in reranker: score = the candidate's own (first-seen) score, best top_k, stable
            order, final = boost_and_order(cand_sc.tolist(), [None] * cn, None, top_k=p["top_k_rerank"])
            rows, sc = cand[order], np.asarray(final, np.float32)
            threshold = p["threshold"]
        else:
            # rerank off: every score 0.0, list cut to top_k_rerank (retriever.py:229-231), no threshold
            rows, sc = cand[: p["top_k_rerank"]], np.zeros(min(cn, p["top_k_rerank"]), np.float32)
            threshold = None
        arts = group_scored_chunks(store, rows[None, :], sc[None, :], threshold=threshold)[0] if len(rows) else []
        got = [{"kb_id": a.kb_id, "rerank_score": a.score, "normalized_rank": a.normalized_rank,
                "article_rank": a.article_rank, "matched": [f"{r:012d}" for r in a.rows]} for a in arts]
        want = case["articles"]
        assert len(got) == len(want), case["name"]
        for g_, w_ in zip(got, want):
            assert (g_["kb_id"], g_["article_rank"], g_["matched"]) == (w_["kb_id"], w_["article_rank"], w_["matched"])
            assert g_["normalized_rank"] == pytest.approx(w_["normalized_rank"], abs=1e-12)
            assert g_["rerank_score"] == pytest.approx(w_["rerank_score"], abs=1e-6)
    store.close()


def test_threshold_cases_mirror_reference_tests(torch_cuda):
    """rag_engine/tests/test_retriever.py:333-462: articles below the threshold are filtered, all filtered -> empty
    list, a score exactly at the threshold is INCLUDED (`>=`)."""
    from cmw_rag_b200.articles import group_scored_chunks, retrieval_confidence
    from cmw_rag_b200.store import B200Store

    store = B200Store(collection_name="thr", capacity=16)
    kb = ["high_score", "low_score", "medium_score", "123", "exact_threshold", "above_threshold"]
    rng = np.random.default_rng(0)
    store.add([f"chunk{i}" for i in range(6)], [{"kbId": kb[i]} for i in range(6)], ids=[f"c{i}" for i in range(6)],
              embeddings=rng.standard_normal((6, 8)).astype(np.float32))
    rows = np.array([[0, 1, 2], [3, -1, -1], [4, 5, -1]], np.int64)
    sc = np.array([[0.9, 0.1, 0.5], [0.1, -np.inf, -np.inf], [0.5, 0.51, -np.inf]], np.float32)
    arts = group_scored_chunks(store, rows, sc, threshold=0.5)
    assert {a.kb_id for a in arts[0]} == {"high_score", "medium_score"} and len(arts[0]) == 2
    assert arts[1] == []
    assert [a.kb_id for a in arts[2]] == ["above_threshold", "exact_threshold"]
    assert [a.article_rank for a in arts[2]] == [0, 1] and [a.normalized_rank for a in arts[2]] == [0.0, 1.0]
    conf = retrieval_confidence(sc[0])
    assert conf["top_score"] == pytest.approx(0.9) and conf["n_above_threshold"] == 2 and conf["likely_relevant"]
    store.close()


def test_concurrent_callers_through_the_seam_on_the_gpu(cfg1):
    """48 threads and 32 asyncio tasks hit one B200Store at once: every caller gets the oracle's answer for ITS
    query, the launches are shared (batch sizes > 1 in the histogram), and the seam latency stays bounded by a
    few launches -- not by one launch per caller."""
    import asyncio
    import threading
    import time

    from cmw_rag_b200.store import B200Store

    c, q = cfg1["c"], cfg1["q"]
    store = B200Store(collection_name="conc", capacity=c.shape[0], batch_window_us=300, max_batch=32)
    store.add([f"t{i}" for i in range(c.shape[0])], [{"kbId": str(i // 8), "stable_id": f"{i:06d}"} for i in range(c.shape[0])],
              ids=[f"{i:06d}" for i in range(c.shape[0])], embeddings=c)
    store.similarity_search(q[0], k=5)  # warm-up (first launch allocates workspaces)
    ref = cfg1["ref_ids"]
    out, lat = {}, {}

    def worker(i):
        t0 = time.perf_counter()
        out[i] = [d_.metadata["stable_id"] for d_ in store.similarity_search(q[i % 64], k=20)]
        lat[i] = time.perf_counter() - t0

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(48)]
    t_all = time.perf_counter()
    for t in threads:
        t.start()

    async def main():
        return await asyncio.gather(*[store.similarity_search_async(q[i].tolist(), k=20) for i in range(32)])

    res = asyncio.run(main())
    for t in threads:
        t.join()
    wall = time.perf_counter() - t_all
    for i in range(48):
        assert out[i] == [f"{r:06d}" for r in ref[i % 64]], i
    for i, r in enumerate(res):
        assert [d_.metadata["stable_id"] for d_ in r] == [f"{x:06d}" for x in ref[i]]
    m = store.metrics()
    assert m["counters"]["requests"] == 81 and m["counters"]["launches"] <= 20, m["counters"]
    assert m["batch_size"]["max"] >= 8 and m["counters"]["flagged"] == 0
    # 80 callers in a handful of launches: far below 80 sequential single-query searches
    single = m["launch_ms"]["mean"]
    print(f"80 concurrent callers: {m['counters']['launches'] - 1} launches, wall {wall * 1e3:.1f} ms, seam latency p50 "
          f"{m['seam_latency_ms']['p50']} ms p99 {m['seam_latency_ms']['p99']} ms, launch mean {single:.3f} ms")
    assert max(lat.values()) < 0.5
    store.close()


def test_auto_compaction_and_device_side_growth(torch_cuda):
    """Store maintenance without a host round trip (cmw_store_copy_rows): a collection that outgrows its HBM
    reservation is re-ingested device to device, and auto_compact reclaims tombstones once they dominate."""
    from cmw_rag_b200.store import B200Store

    n, d = 9000, 64
    c = synth.make_corpus(n, d, seed=21, ties=False)
    store = B200Store(collection_name="maint", capacity=4096, auto_compact=0.5)
    for lo in range(0, n, 3000):  # 4096 -> 8192 -> 16384 rows reserved: two device-side growths
        store.add([f"t{i}" for i in range(lo, lo + 3000)], [{"kbId": str(i // 4), "doc_stable_id": f"d{i // 100}"} for i in range(lo, lo + 3000)],
                  ids=[f"{i:05d}" for i in range(lo, lo + 3000)], embeddings=c[lo:lo + 3000])
    assert store.dense.capacity >= n and store.dense.rows == n
    q, _ = synth.make_queries(c, 8, seed=3, tie_probe=False)
    ref, _, _ = exact_topk_c(c, q, 5)
    _, ids, fl = store.search(q, 5)
    assert (ids == ref).all() and (fl == 0).all()
    # delete 50 of the 90 documents one by one: the 46th delete crosses auto_compact = 0.5 (4600 of 9000 rows dead)
    # -> physical compaction to 4400 rows, renumbered; the last four deletes are tombstones again (400 of 4400)
    for doc in range(0, 50):
        store.delete(where={"doc_stable_id": f"d{doc}"})
    assert store.count() == n - 5000 and store.dense.rows == 4400, (store.count(), store.dense.rows)
    assert store.dense.live_rows == 4000
    live = np.zeros(n, np.uint8)
    live[5000:] = 1
    ref2, _, _ = exact_topk_c(c, q, 5, live=live)
    got = [[d_.metadata["doc_stable_id"] for d_ in store.similarity_search(q[b], k=5)] for b in range(8)]
    assert got == [[f"d{r // 100}" for r in ref2[b]] for b in range(8)]
    store.close()


def test_tf32_filter_over_fp32_tiles(cfg1, torch_cuda):
    """A store WITHOUT 16-bit tiles filters through kind::tf32 MMAs over the fp32 rows: one pass for any batch size
    (K1 needs one pass per 4 queries).  Device API: the oracle's ids, nothing flagged; also selectable on a store
    that has 16-bit tiles (algo="gemm_tf32"); and the D = 200 / 768 dims exercise the partial last k-block."""
    torch = torch_cuda
    from cmw_rag_b200 import DenseStore

    c, q = cfg1["c"], cfg1["q"]
    st = DenseStore(1536, c.shape[0], bf16=False)
    st.append(c)
    assert st.info()["gemm_ready"]
    qd = torch.from_numpy(q).cuda()
    for b in (1, 16, 40, 64):
        sc, ids, fl = st.search(qd[:b], 20, mode="f32")  # auto -> tf32 filter
        torch.cuda.synchronize()
        _check_exact(ids.cpu().numpy(), sc.cpu().numpy(), cfg1["ref_ids"][:b], cfg1["ref_sc"][:b])
        assert int(fl.sum()) == 0, b
    sc, ids, fl = cfg1["store"].search(qd, 20, mode="f32", algo="gemm_tf32")
    torch.cuda.synchronize()
    _check_exact(ids.cpu().numpy(), sc.cpu().numpy(), cfg1["ref_ids"], cfg1["ref_sc"])
    assert int(fl.sum()) == 0
    sc, ids, fl = st.search_host(q[:9], 20, metric="ip", mode="f32")
    ref_ids, ref_sc, _ = exact_topk_c(c, q[:9], 20, metric="ip")
    _check_exact(ids, sc, ref_ids, ref_sc)
    st.close()
    for n, d in ((30000, 200), (20011, 768)):
        c2 = synth.make_corpus(n, d, seed=n, ties=False)
        q2, _ = synth.make_queries(c2, 33, seed=2, tie_probe=False)
        s2 = DenseStore(d, n, bf16=False)
        s2.append(c2)
        ref_ids, ref_sc, _ = exact_topk_c(c2, q2, 10)
        sc, ids, fl = s2.search(torch.from_numpy(q2).cuda(), 10, mode="f32")
        torch.cuda.synchronize()
        _check_exact(ids.cpu().numpy(), sc.cpu().numpy(), ref_ids, ref_sc)
        assert int(fl.sum()) == 0, (n, d)
        s2.close()
