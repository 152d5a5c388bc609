"""CPU-side tests (no GPU): the C-ABI library loads and exports every symbol the header declares,
the compute entry points fail loudly without a device, and the host logic of the Python layer
(kbId keys, where-filters, sidecar bookkeeping, await coalescing, shard plumbing) behaves like the
reference's store.  Where a search result is needed the oracle is injected as the backend -- test
infrastructure only; the product classes have no CPU path of their own."""
from __future__ import annotations

import asyncio
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import synth
from oracle import exact_topk, merge_topk as oracle_merge
from oracle.multivector import extract_numeric_kbid as oracle_kbid, group_key as oracle_group_key

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="session")
def native():
    from cmw_rag_b200 import build

    build.build()
    from cmw_rag_b200 import _native

    return _native


def test_library_exports_every_declared_symbol(native):
    with open(os.path.join(ROOT, "include", "cmw_dense.h")) as f:
        header = f.read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(cmw_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 18
    lib = ctypes.CDLL(native.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/cmw_dense.h but not exported"
    assert declared == set(native.SIGNATURES), declared ^ set(native.SIGNATURES)
    assert native.lib().cmw_abi_version() == int(re.search(r"CMW_ABI_VERSION (\d+)", header).group(1))


def test_no_cpu_fallback(native):
    """Without a CUDA device the store cannot be created and says why."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    from cmw_rag_b200 import DenseStore

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        DenseStore(1536, 16)
    assert native.kernel_launches() == 0


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under cmw_rag_b200/ may reference it."""
    pkg = os.path.join(ROOT, "cmw_rag_b200")
    for dirpath, _, files in os.walk(pkg):
        if os.path.basename(dirpath) == "build":
            continue
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                with open(os.path.join(dirpath, fn)) as f:
                    src = f.read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn
                assert "liboracle" not in src and "oracle_topk" not in src and "oracle." not in src, fn


def test_kbid_keys_match_reference_behaviour():
    from cmw_rag_b200 import extract_numeric_kbid, group_key

    for v in ["4578-toc", "abc", 12, None, "", "12ab34", "007", "-5", " 12", 0, "٣٤"]:
        assert extract_numeric_kbid(v) == oracle_kbid(v), v
        assert group_key(v) == oracle_group_key(v), v


def test_where_filter():
    from cmw_rag_b200.store import _match

    meta = {"kbId": "12", "n": 3, "flag": True}
    assert _match(meta, None) and _match(meta, {})
    assert _match(meta, {"kbId": "12"}) and not _match(meta, {"kbId": 12})
    assert not _match(meta, {"missing": 1})
    assert _match(meta, {"n": {"$eq": 3}}) and not _match(meta, {"n": {"$ne": 3}})
    assert _match(meta, {"n": {"$in": [1, 3]}}) and _match(meta, {"n": {"$nin": [1, 2]}})
    assert _match(meta, {"$and": [{"kbId": "12"}, {"n": 3}]})
    assert _match(meta, {"$or": [{"kbId": "x"}, {"flag": True}]})
    assert not _match(meta, {"$and": [{"kbId": "12"}, {"n": 4}]})


class FakeDense:
    """Stands in for DenseStore in host-logic tests: answers with the oracle."""

    def __init__(self, dim, capacity, device=0, f32=True, bf16=True, id_offset=0, tiles16="f16"):
        self.dim, self.id_offset = dim, id_offset
        self.rows_ = np.zeros((0, dim), np.float32)
        self.live = np.zeros((0,), bool)
        self.gid = np.zeros((0,), np.int32)
        self.calls = []

    def append(self, rows, kb_gid=None):
        rows = np.asarray(rows, np.float32)
        self.rows_ = np.concatenate([self.rows_, rows])
        self.live = np.concatenate([self.live, np.ones(len(rows), bool)])
        self.gid = np.concatenate([self.gid, np.asarray(kb_gid, np.int32) if kb_gid is not None
                                   else np.full(len(rows), -1, np.int32)])

    def tombstone(self, rows):
        self.live[np.asarray(rows, np.int64)] = False

    def close(self):
        pass

    def read_rows(self, row0, n, rows=True):
        return self.rows_[row0:row0 + n].copy(), self.gid[row0:row0 + n].copy(), self.live[row0:row0 + n].copy()

    def copy_rows_from(self, src, rows=None, row0=0, n=None):
        idx = np.asarray(rows, np.int64) if rows is not None else np.arange(row0, row0 + (len(src.rows_) - row0 if n is None else n))
        self.append(src.rows_[idx], src.gid[idx])

    @property
    def rows(self):
        return len(self.rows_)

    def search_host(self, q, k, metric="cosine", mode="f32", algo=None):
        self.calls.append(q.shape[0])
        ids, sc, _ = exact_topk(self.rows_, q, k, metric=metric, live=self.live, id_offset=self.id_offset)
        return sc, ids, np.zeros(q.shape[0], np.int32)


@pytest.fixture()
def fake_store(monkeypatch):
    import cmw_rag_b200.store as store_mod

    monkeypatch.setattr(store_mod, "DenseStore", FakeDense)
    return store_mod.B200Store(collection_name="t", capacity=1024)


def test_store_mirrors_reference_store_tests(fake_store):
    """rag_engine/tests/test_storage_vector_store.py:10-70 of the reference, against B200Store."""
    store = fake_store

    async def run():
        await store.add_async(texts=["a", "b"], metadatas=[{"kbId": "doc1"}, {"kbId": "doc2"}],
                              ids=["1", "2"], embeddings=[[0.1, 0.0, 0.0], [0.0, 0.1, 0.0]])
        results = await store.similarity_search_async(query_embedding=[0.1, 0.0, 0.0], k=1)
        assert len(results) == 1 and results[0].metadata["kbId"] == "doc1"
        assert results[0].page_content == "a"
        await store.add_async(texts=["t"], metadatas=[{"doc_stable_id": "D", "kbId": "k9", "file_mtime_epoch": 5}],
                              ids=["x"], embeddings=[[0.0, 0.0, 1.0]])
        meta = await store.get_any_doc_meta_async({"doc_stable_id": "D"})
        assert meta["file_mtime_epoch"] == 5
        assert (await store.get_by_kb_id_async("k9"))["doc_stable_id"] == "D"
        assert await store.get_by_kb_id_async("nope") is None
        await store.delete_where_async({"doc_stable_id": "D"})
        assert await store.get_any_doc_meta_async({"doc_stable_id": "D"}) is None
        assert store.count() == 2
        # a deleted row is never returned, and k larger than the collection returns fewer results
        res = await store.similarity_search_async(query_embedding=[0.0, 0.0, 1.0], k=5)
        assert len(res) == 2
        # adding a known id again is ignored (Chroma add semantics), new ids are appended
        await store.add_async(texts=["dup", "new"], metadatas=[{"kbId": "zz"}, {"kbId": "doc3-toc"}],
                              ids=["1", "3"], embeddings=[[1.0, 0, 0], [0, 0, 1.0]])
        assert store.count() == 3
        got = await store.similarity_search_async(query_embedding=[0.0, 0.0, 1.0], k=1)
        assert got[0].metadata["kbId"] == "doc3-toc"
        col = await store.get_collection()
        raw = await col.query(query_embeddings=[[0.1, 0.0, 0.0]], n_results=2, include=["documents", "metadatas"])
        assert raw["ids"] == [["1", "2"]] and raw["documents"] == [["a", "b"]] and raw["distances"] is None
        assert await col.count() == 3
        with pytest.raises(ValueError):
            await store.add_async(texts=["q"], metadatas=[{}], ids=["9"], embeddings=[[1.0, 2.0]])  # wrong dim

    asyncio.run(run())
    assert store.gid_for("doc3") == store.gid_for("doc3-toc") or store.gid_for("doc3-toc") >= 0
    assert store.gid_for("") == -1 and store.gid_for(None) == -1
    assert store.gid_for("4578-toc") == store.gid_for("4578")


def test_seam_and_await_coalescing(fake_store):
    """rag_engine/tests/test_retrieval_vector_search.py:11-20 (kwargs passed literally) and the
    fan-out of retriever.py:179-182: S concurrent awaits become ONE batched search."""
    from cmw_rag_b200 import top_k_search_async

    store = fake_store
    rng = np.random.default_rng(0)
    emb = rng.standard_normal((50, 8)).astype(np.float32)
    store.add([f"t{i}" for i in range(50)], [{"stable_id": f"{i:04d}", "kbId": str(100 + i // 5)} for i in range(50)],
              ids=[f"{i:04d}" for i in range(50)], embeddings=emb)
    qs = rng.standard_normal((4, 8)).astype(np.float32)

    async def run():
        outs = await asyncio.gather(*[top_k_search_async(store, q.tolist(), k=3 + i) for i, q in enumerate(qs)])
        return outs

    outs = asyncio.run(run())
    assert store.dense.calls == [4]  # one launch for the four awaits
    ids, _, _ = exact_topk(emb, qs, 6)
    for i, docs in enumerate(outs):
        assert [d.metadata["stable_id"] for d in docs] == [f"{r:04d}" for r in ids[i, : 3 + i]]
    assert store.stats["max_batch"] == 4 and store.stats["launch_batches"] == 1

    class Recorder:
        def __init__(self):
            self.kw = None

        async def similarity_search_async(self, **kw):
            self.kw = kw
            return ["x"]

    rec = Recorder()
    assert asyncio.run(top_k_search_async(rec, [0.1, 0.2], 3)) == ["x"]
    assert rec.kw == {"query_embedding": [0.1, 0.2], "k": 3}


def test_save_load_roundtrip_host_layer(fake_store, tmp_path):
    """the on-disk format restores rows, tombstones, sidecar and the kbId group table."""
    import cmw_rag_b200.store as store_mod

    store = fake_store
    rng = np.random.default_rng(5)
    emb = rng.standard_normal((40, 12)).astype(np.float32)  # 12 -> padded to 16 internally
    metas = [{"stable_id": f"s{i}", "kbId": f"{100 + i // 4}" + ("-toc" if i % 7 == 0 else ""), "n": i} for i in range(40)]
    store.add([f"doc {i}" for i in range(40)], metas, ids=[f"s{i}" for i in range(40)], embeddings=emb)
    store.delete(where={"n": {"$in": [3, 17]}})
    store.save(str(tmp_path / "col"))
    again = store_mod.B200Store.load(str(tmp_path / "col"))
    assert again.count() == 38 and again.collection_name == "t"
    q = rng.standard_normal((5, 12)).astype(np.float32)
    s1, i1, _ = store.search(q, 7)
    s2, i2, _ = again.search(q, 7)
    assert (i1 == i2).all() and np.array_equal(s1, s2)
    assert [d.metadata for d in again.similarity_search(q[0], 5)] == [d.metadata for d in store.similarity_search(q[0], 5)]
    assert again.get(where={"n": 3})["ids"] == [] and again.get(where={"n": 4})["ids"] == ["s4"]
    assert again.gid_for("100") == store.gid_for("100-toc")
    again.add(["new"], [{"kbId": "999"}], ids=["new"], embeddings=rng.standard_normal((1, 12)))
    assert again.count() == 39


def test_collection_grows_past_its_capacity(monkeypatch):
    import cmw_rag_b200.store as store_mod

    monkeypatch.setattr(store_mod, "DenseStore", FakeDense)
    store = store_mod.B200Store(collection_name="g", capacity=8)
    rng = np.random.default_rng(1)
    emb = rng.standard_normal((30, 8)).astype(np.float32)
    store.add([f"a{i}" for i in range(6)], [{"kbId": str(i)} for i in range(6)], ids=[f"a{i}" for i in range(6)], embeddings=emb[:6])
    store.delete(ids=["a2"])
    store.add([f"b{i}" for i in range(24)], [{"kbId": str(100 + i)} for i in range(24)], ids=[f"b{i}" for i in range(24)],
              embeddings=emb[6:])
    assert store.count() == 29 and store._capacity >= 30
    live = np.ones(30, bool)
    live[2] = False
    ref, _, _ = exact_topk(emb, emb[:4], 5, live=live)
    _, ids, _ = store.search(emb[:4], 5)
    assert (ids == ref).all()


def test_where_index_and_compaction(fake_store):
    """the indexer's per-document lookups (`get_any_doc_meta_async({"doc_stable_id": ..})`,
    `delete_where_async`, core/indexer.py:397,432,505) hit a hash index instead of scanning the sidecar, and
    compact() reclaims tombstoned rows without changing what a search returns."""
    store = fake_store
    rng = np.random.default_rng(3)
    n = 400
    emb = rng.standard_normal((n, 8)).astype(np.float32)
    metas = [{"doc_stable_id": f"D{i // 4}", "stable_id": f"c{i}", "kbId": str(500 + i // 4), "n": i} for i in range(n)]
    store.add([f"t{i}" for i in range(n)], metas, ids=[f"c{i}" for i in range(n)], embeddings=emb)
    assert store._index_candidates({"doc_stable_id": "D7"}) == [28, 29, 30, 31]
    assert store._index_candidates({"n": 5}) is None  # not an indexed key: full scan
    assert store._index_candidates({"doc_stable_id": {"$eq": "D7"}, "n": 30}) == [28, 29, 30, 31]
    assert store.get(where={"doc_stable_id": "D7", "n": 30})["ids"] == ["c30"]
    assert store.get(where={"doc_stable_id": {"$ne": "D7"}}, limit=2)["ids"] == ["c0", "c1"]  # scan path
    assert store.get(where={"kbId": "507"})["ids"] == ["c28", "c29", "c30", "c31"]
    # re-index documents 0..49: delete their chunks, add new versions (the indexer's update flow)
    for doc in range(50):
        assert store.delete(where={"doc_stable_id": f"D{doc}"}) == 4
    assert store.count() == n - 200 and store.get(where={"doc_stable_id": "D7"})["ids"] == []
    emb2 = rng.standard_normal((200, 8)).astype(np.float32)
    store.add([f"u{i}" for i in range(200)],
              [{"doc_stable_id": f"D{i // 4}", "stable_id": f"v{i}", "kbId": str(500 + i // 4)} for i in range(200)],
              ids=[f"v{i}" for i in range(200)], embeddings=emb2)
    assert store.get(where={"doc_stable_id": "D7"})["ids"] == ["v28", "v29", "v30", "v31"]
    q = rng.standard_normal((6, 8)).astype(np.float32)
    before = [[d.metadata["stable_id"] for d in store.similarity_search(v, 9)] for v in q]
    dist_before = store.query(q, 9)["distances"]
    assert store.compact() == 200
    assert store.count() == n and len(store._ids) == n and store.compact() == 0
    after = [[d.metadata["stable_id"] for d in store.similarity_search(v, 9)] for v in q]
    assert after == before and store.query(q, 9)["distances"] == dist_before
    assert store.get(where={"doc_stable_id": "D7"})["ids"] == ["v28", "v29", "v30", "v31"]
    assert store.get(ids=["c399", "c0"])["ids"] == ["c399"]
    _, ids, _ = store.search(q[:1], 3)
    assert all(store._ids[r] == sid for r, sid in zip(ids[0].tolist(), before[0][:3]))
    store.add(["w"], [{"doc_stable_id": "Dnew", "kbId": "9"}], ids=["w"], embeddings=rng.standard_normal((1, 8)))
    assert store.get(where={"doc_stable_id": "Dnew"})["ids"] == ["w"] and store.count() == n + 1


def test_shard_bounds():
    from cmw_rag_b200.sharded import shard_bounds

    assert shard_bounds(10, 1) == [(0, 10)]
    assert shard_bounds(10, 4) == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert shard_bounds(3, 8)[3:] == [(3, 3)] * 5
    b = shard_bounds(200_000_000, 8)
    assert b[0] == (0, 25_000_000) and b[-1] == (175_000_000, 200_000_000)


_GLOO_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import numpy as np, torch, torch.distributed as dist
import synth
from oracle import exact_topk, merge_topk
from cmw_rag_b200.sharded import ShardedSearcher, shard_bounds
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n, d, k = 1003, 32, 10
c = synth.make_corpus(n, d, seed=1)
q, _ = synth.make_queries(c, 7, seed=2)
lo, hi = shard_bounds(n, world)[rank]
def local_search(qt, k, **kw):
    ids, sc, sc64 = exact_topk(c[lo:hi], qt.numpy(), k, id_offset=lo)
    return torch.from_numpy(sc64), torch.from_numpy(ids), torch.zeros(qt.shape[0], dtype=torch.int32)
def merge(s64, ids, k):
    mi, ms = merge_topk(ids.numpy(), s64.numpy(), k)
    return torch.from_numpy(ms), torch.from_numpy(mi), None
s = ShardedSearcher(local_search=local_search, merge=merge)
ms, mi, fl = s.search(torch.from_numpy(q), k)
gi, gs, _ = exact_topk(c, q, k)
assert (mi.numpy() == gi).all(), (rank, mi, gi)
assert np.abs(ms.numpy() - gs).max() == 0.0
# a world larger than the corpus: empty shards contribute only padding
tiny = c[:1]
lo2, hi2 = shard_bounds(1, world)[rank]
def ls2(qt, k, **kw):
    ids, sc, sc64 = exact_topk(tiny[lo2:hi2], qt.numpy(), k, id_offset=lo2) if hi2 > lo2 else (
        np.full((qt.shape[0], k), -1, np.int64), None, np.full((qt.shape[0], k), -np.inf))
    return torch.from_numpy(sc64), torch.from_numpy(ids), None
ms, mi, _ = ShardedSearcher(local_search=ls2, merge=merge).search(torch.from_numpy(q), 3)
assert (mi.numpy()[:, 0] == 0).all() and (mi.numpy()[:, 1:] == -1).all()
dist.barrier(); dist.destroy_process_group()
print("ok", rank)
"""


def test_sharded_search_world_size_2_gloo(tmp_path):
    """The N > 1 path under gloo on CPU: shard bounds, id offsets, all-gather layout, merge."""
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="2")
    res = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
         "--master-addr", "127.0.0.1", "--master-port", "29541", str(script)],
        capture_output=True, text=True, timeout=300, env=env)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert res.stdout.count("ok") == 2


_GLOO_TWO_PHASE_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import numpy as np, torch, torch.distributed as dist
import synth
from oracle import exact_topk, exact_scores, merge_topk
from cmw_rag_b200.sharded import ShardedSearcher, shard_bounds
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n, d, k = 1501, 32, 12
c = synth.make_corpus(n, d, seed=1)
q, _ = synth.make_queries(c, 9, seed=2)
lo, hi = shard_bounds(n, world)[rank]
EPS = 1e-3

class OracleBackend:
    # the four device steps of the two-phase search, restated with the oracle (test stand-in for the kernels)
    def __init__(self):
        self.rescored = 0
    def filter(self, qt, k, **kw):
        self.s = exact_scores(c[lo:hi], qt.numpy())                       # [B, n_local] fp64
        top = -np.sort(-self.s, axis=1)[:, :k]
        return torch.from_numpy(top.astype(np.float32))
    def kth(self, gathered, k):
        g = gathered.numpy()                                               # [G, B, k]
        assert g.shape == (world, q.shape[0], k)
        flat = np.transpose(g, (1, 0, 2)).reshape(g.shape[1], -1)
        return torch.from_numpy(-np.sort(-flat, axis=1)[:, k - 1].copy())
    def finish(self, qt, k, kth, **kw):
        b = qt.shape[0]
        block = np.full((b, 2 * k + 1), -np.inf)
        ids = np.full((b, k), -1, np.int64)
        for i in range(b):
            cut = float(kth[i]) - 2 * EPS if kth is not None else -np.inf
            cand = np.flatnonzero(self.s[i] >= cut)
            self.rescored += cand.size
            order = cand[np.lexsort((cand, -self.s[i][cand]))][:k]
            block[i, : order.size] = self.s[i][order]
            ids[i, : order.size] = order + lo
        block[:, k:2 * k] = ids.view(np.float64)
        block[:, 2 * k] = 0.0
        return torch.from_numpy(block.reshape(-1))
    def merge(self, blocks, w, b, k):
        g = blocks.numpy().reshape(w, b, 2 * k + 1)
        mi, ms = merge_topk(np.ascontiguousarray(g[:, :, k:2 * k]).view(np.int64), np.ascontiguousarray(g[:, :, :k]), k)
        return torch.from_numpy(ms), torch.from_numpy(mi), torch.zeros(b, dtype=torch.int32)

be = OracleBackend()
s = ShardedSearcher(backend=be)
ms, mi, fl = s.search(torch.from_numpy(q), k)
gi, gs, _ = exact_topk(c, q, k)
assert (mi.numpy() == gi).all(), (rank, mi, gi)
assert np.abs(ms.numpy() - gs).max() == 0.0
# the point of the exchange: the shards together rescore little more than ONE shard's worth of candidates
t = torch.tensor([be.rescored], dtype=torch.int64); dist.all_reduce(t)
assert int(t) < 3 * k * q.shape[0], int(t)
dist.barrier(); dist.destroy_process_group()
print("ok", rank)
"""


def test_sharded_two_phase_search_world_size_2_gloo(tmp_path):
    """Host logic of the two-phase row-sharded search (filter | all-gather | global k-th | finish | all-gather |
    merge) under gloo on CPU, the oracle standing in for the four device steps."""
    script = tmp_path / "worker2.py"
    script.write_text(_GLOO_TWO_PHASE_WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="2")
    res = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
         "--master-addr", "127.0.0.1", "--master-port", "29543", str(script)],
        capture_output=True, text=True, timeout=300, env=env)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert res.stdout.count("ok") == 2


def test_reference_arm_json_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) on a tiny corpus: one JSON line with
    the contract's keys, rank 0 only."""
    import json

    env = dict(os.environ, RANK="0", WORLD_SIZE="1")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--rows", "12000",
                          "--hnsw-rows", "12000", "--steps", "2", "--warmup", "1", "--hnsw-queries", "32", "--k", "10"],
                         capture_output=True, text=True, timeout=600, env=env)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "queries/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 2 and d["warmup"] == 1 and d["vs_baseline"] is None
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "HNSW" in cb["sample"]
    assert cb["index_rows"] == 12000 and 0.0 <= cb["recall_at_k"] <= 1.0
    # the other ranks of a torchrun launch exit without work and without output
    res2 = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                          capture_output=True, text=True, timeout=120, env=dict(os.environ, RANK="1", WORLD_SIZE="2"))
    assert res2.returncode == 0 and res2.stdout.strip() == ""


def test_committed_bench_line_has_the_contract_keys():
    """The last bench line of the round (profiles/) carries every key of the driver contract."""
    import json

    with open(os.path.join(ROOT, "profiles", "r01z_bench_final.json")) as f:
        d = json.loads(f.read().strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert key in d, key
    assert "workload" in d["config"] and "model" not in d["config"]
    assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(d["e2e"])
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(d["roofline"])
    assert set(("value", "unit", "cores", "kind", "sample")) <= set(d["cpu_baseline"])
    assert set(("sm_mhz", "sm_max_mhz", "reasons")) <= set(d["clocks"])
    assert d["gpu_launches"] > 0 and d["parity"]["ids_identical_to_fp64_oracle"] is True


def test_scan_order_permutation_is_a_bijection_and_spreads_evenly():
    """The arithmetic of K2's scan order (set_scan_order in csrc/gemm.cu, scan_tile in csrc/gemm_common.cuh),
    restated: tile g of the scan is store tile (g * m) mod T with m ~ T / golden ratio, coprime to T.  Every tile is
    visited exactly once, and any range of consecutive scan positions is spread over the whole store."""
    from math import gcd

    for T in (64, 65, 100, 511, 512, 3907, 7813, 97657, 781250):
        m = int(T * 0.6180339887498949) | 1
        while gcd(m, T) != 1:
            m += 2
        m %= T
        g = np.arange(T, dtype=np.int64)
        perm = (g * m) % T
        assert np.array_equal(np.sort(perm), g), T
        # a slab of 1/16 of the scan positions, anywhere: its tiles leave no gap wider than a few times the ideal
        # spacing (1.2-4.8x for these sizes)
        n = max(8, T // 16)
        for start in (0, T // 3, T - n):
            tiles = np.sort(perm[start:start + n])
            gaps = np.diff(np.concatenate([tiles, [tiles[0] + T]]))
            assert gaps.max() <= 8 * (T / n) + 2, (T, start, int(gaps.max()))


# ------------------------------------------------------------------------------------------------
# round 2: rerank-side statistics, batching front-end, versioned collections
# ------------------------------------------------------------------------------------------------
def test_confidence_statistics_match_reference_golden(golden_dir):
    """retrieval/confidence.py:13-117 of the reference, through golden vectors its own functions produced
    (tests/golden/make_golden_f3.py)."""
    import json

    from cmw_rag_b200.articles import (normalized_confidence_from_traces, retrieval_confidence,
                                       retrieval_confidence_batch)

    with open(os.path.join(golden_dir, "f3_golden.json")) as f:
        g = json.load(f)
    assert len(g["confidence"]) >= 50
    for case in g["confidence"]:
        got = retrieval_confidence(case["scores"], case["threshold"], case["mean_top_k"])
        want = case["expected"]
        assert got["n_above_threshold"] == want["n_above_threshold"] and got["likely_relevant"] == want["likely_relevant"]
        for key in ("top_score", "mean_top_k", "score_gap"):
            assert got[key] == pytest.approx(want[key], abs=1e-12), (key, case)
    for case in g["normalized"]:
        got = normalized_confidence_from_traces(case["traces"])
        assert (got is None) == (case["expected"] is None)
        if got is not None:
            assert got == pytest.approx(case["expected"], abs=1e-12)
    rows = np.array([[0.9, 0.2, 0.6, 0.0], [0.1, 0.1, 0.0, 0.0]])
    out = retrieval_confidence_batch(rows, counts=[3, 2])
    assert out[0] == retrieval_confidence([0.9, 0.2, 0.6]) and out[1] == retrieval_confidence([0.1, 0.1])


def test_batcher_window_batchsize_backpressure_and_metrics():
    import threading
    import time

    from cmw_rag_b200.batcher import QueueFull, SearchBatcher

    calls = []
    gate = threading.Event()

    def search(q, kmax):
        calls.append(q.shape[0])
        gate.wait(5)  # the "GPU" is busy until the test says so
        ids = np.tile(np.arange(kmax), (q.shape[0], 1)) + q[:, :1].astype(np.int64) * 1000
        return np.zeros((q.shape[0], kmax), np.float32), ids, np.zeros(q.shape[0], np.int32)

    b = SearchBatcher(search, max_batch=8, max_wait_us=20_000, max_queue=12, name="t")
    try:
        first = b.submit(np.full(4, 1.0), 3)
        time.sleep(0.08)  # the window (20 ms) expires: a batch of one is dispatched and blocks on the gate
        assert calls == [1]
        # while that launch runs, requests pile up: 12 fit, the 13th is refused without blocking
        futs = [b.submit(np.full(4, float(i + 2)), 2 + i % 3) for i in range(12)]
        with pytest.raises(QueueFull):
            b.submit(np.zeros(4), 1, block=False)
        gate.set()
        sc, ids, flag = first.result(5)
        assert ids.tolist() == [1000, 1001, 1002] and flag == 0
        for i, f in enumerate(futs):
            sc, ids, flag = f.result(5)
            assert ids.shape == (2 + i % 3,) and ids[0] == (i + 2) * 1000  # every caller gets ITS row, cut to ITS k
        assert calls[1] == 8 and sum(calls) == 13  # max_batch caps a launch; nothing is lost
        m = b.metrics()
        assert m["counters"]["requests"] == 13 and m["counters"]["rejected"] == 1
        assert m["batch_size"]["max"] == 8 and m["queue_depth_at_dispatch"]["max"] == 12
        assert m["seam_latency_ms"]["count"] == 13 and m["queue_wait_ms"]["p99"] > 0
        text = b.prometheus()
        assert "cmw_t_batch_size_bucket" in text and "cmw_t_requests_total 13" in text
        # an exception in the launch reaches every waiter of that batch
        b._search = lambda q, k: (_ for _ in ()).throw(RuntimeError("boom"))
        f = b.submit(np.zeros(4), 1)
        with pytest.raises(RuntimeError, match="boom"):
            f.result(5)
    finally:
        gate.set()
        b.close()


def test_concurrent_callers_share_launches_through_the_store(fake_store):
    """Threads calling similarity_search and asyncio tasks awaiting similarity_search_async at the same time end up
    in shared launches (the reference issues one HTTP query per call: retriever.py:179-182,
    retrieve_context.py:397-409), each getting its own answer."""
    import asyncio
    import threading

    store = fake_store
    rng = np.random.default_rng(5)
    n, d = 300, 8
    emb = rng.standard_normal((n, d)).astype(np.float32)
    store.add([f"t{i}" for i in range(n)], [{"kbId": str(i // 3), "stable_id": f"s{i}"} for i in range(n)],
              ids=[f"s{i}" for i in range(n)], embeddings=emb)
    out = {}

    def worker(i):
        out[i] = [d_.metadata["stable_id"] for d_ in store.similarity_search(emb[i] * 1.5, k=1)]

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(24)]
    for t in threads:
        t.start()

    async def main():
        return await asyncio.gather(*[store.similarity_search_async(emb[100 + i].tolist(), k=2) for i in range(12)])

    res = asyncio.run(main())
    for t in threads:
        t.join()
    assert all(out[i] == [f"s{i}"] for i in range(24))
    assert all(r[0].metadata["stable_id"] == f"s{100 + i}" and len(r) == 2 for i, r in enumerate(res))
    st = store.stats
    assert st["searches"] == 36 and st["launch_batches"] < 36 and st["max_batch"] >= 12
    m = store.metrics()
    assert m["collection"]["live_rows"] == n and m["batch_size"]["count"] == st["launch_batches"]
    store.close()


def test_collection_registry_names_cache_and_persistence(monkeypatch, tmp_path):
    """config/settings.py:261-273 (collection name per product version) and the per-version store cache of
    tools/retrieve_context.py:101-131."""
    import cmw_rag_b200.store as store_mod
    from cmw_rag_b200.registry import CollectionRegistry

    monkeypatch.setattr(store_mod, "DenseStore", FakeDense)
    reg = CollectionRegistry("kb", overrides={"v5": "legacy_five", "v6": ""}, root=str(tmp_path), capacity=64)
    # the reference's three rules: override if non-empty, else {base}_{version}, unknown version -> base
    assert reg.collection_name("v5") == "legacy_five"
    assert reg.collection_name("v6") == "kb_v6"
    assert reg.collection_name("v7") == "kb" and reg.collection_name(None) == "kb"
    s5, s6 = reg.get_store("v5"), reg.get_store("v6")
    assert s5 is reg.get_store("v5") and s5 is not s6 and s5.collection_name == "legacy_five"
    rng = np.random.default_rng(1)
    e5 = rng.standard_normal((10, 8)).astype(np.float32)
    s5.add([f"five{i}" for i in range(10)], [{"kbId": str(i)} for i in range(10)], ids=[f"a{i}" for i in range(10)],
           embeddings=e5)
    s6.add(["six"], [{"kbId": "600"}], ids=["b0"], embeddings=rng.standard_normal((1, 8)).astype(np.float32))
    assert s5.count() == 10 and s6.count() == 1  # one store per product version
    path = reg.save("v5")
    assert path.endswith("legacy_five") and os.path.exists(os.path.join(path, "meta.json"))
    reg.close()
    reg2 = CollectionRegistry("kb", overrides={"v5": "legacy_five"}, root=str(tmp_path), capacity=64)
    again = reg2.get_store("v5")  # found on disk, like a Chroma collection in the server's --path directory
    assert again.count() == 10 and again.similarity_search(e5[3], k=1)[0].page_content == "five3"
    assert reg2.get_store("v6").count() == 0  # never saved: starts empty
    reg2.close()


def test_upload_slices_partition_the_batch():
    """Row-sharded host form: every rank uploads its own piece of the replicated batch and the NVLink all-gather
    completes it -- the pieces must partition [0, batch) for any batch / world, short and empty pieces included,
    and sit at rank * per in the padded gather buffer."""
    from cmw_rag_b200.sharded import upload_slice

    for world in (1, 2, 3, 4, 7, 8):
        for batch in (1, 2, 7, 8, 9, 37, 64, 1000, 4096, 4097):
            covered = []
            per0 = None
            for rank in range(world):
                lo, hi, per = upload_slice(batch, world, rank)
                per0 = per if per0 is None else per0
                assert per == per0 and per * world >= batch and (per - 1) * world < batch
                assert 0 <= lo <= hi <= batch and hi - lo <= per
                assert lo == min(batch, rank * per)  # the piece starts at its slot of the padded buffer
                covered.extend(range(lo, hi))
            assert covered == list(range(batch)), (world, batch)
