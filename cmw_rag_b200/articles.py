"""Rerank-side score plumbing: what happens to scored chunks after the rerank stage.

Mirrors, for Q queries at once, the tail of ``RAGRetriever.retrieve_async``
(rag_engine/retrieval/retriever.py):

* ``reranker.py:165-181``  metadata boost ``score * (1 + boost)`` and best-first order  (host; the
  cross-encoder itself is out of scope),
* ``retriever.py:234-242``  group chunks by normalised kbId keeping the MAX score -- K4 with one
  "segment" per query: the kernel is score-source-agnostic,
* ``retriever.py:248-260``  inclusive threshold ``score >= rerank_score_threshold``,
* ``retriever.py:307-316``  stable sort by score descending, ``normalized_rank = idx / (n - 1)``.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Sequence

import numpy as np


def metadata_boost(meta: dict[str, Any] | None, weights: dict[str, float] | None) -> float:
    """reranker.py:168-177: additive boosts for tags / code / section heading."""
    boost = 0.0
    if weights and meta:
        if meta.get("tags") and weights.get("tag_match"):
            boost += weights["tag_match"]
        if meta.get("has_code") and weights.get("code_presence"):
            boost += weights["code_presence"]
        if meta.get("section_heading") and weights.get("section_match"):
            boost += weights["section_match"]
    return boost


def boost_and_order(scores: Sequence[float], metas: Sequence[dict | None], weights: dict[str, float] | None,
                    top_k: int | None = None):
    """final = score * (1 + boost); stable best-first order; optional top_k cut (reranker.py:178-181).
    Returns (order indices, final scores in that order)."""
    final = [float(s) * (1.0 + metadata_boost(m, weights)) for s, m in zip(scores, metas)]
    order = sorted(range(len(final)), key=lambda i: final[i], reverse=True)
    if top_k is not None:
        order = order[:top_k]
    return order, [final[i] for i in order]


@dataclass
class ArticleGroup:
    kb_id: str
    score: float
    rows: list[int]                      # member chunk rows in candidate order
    normalized_rank: float = 0.0
    article_rank: int = 0
    metadata: dict[str, Any] = field(default_factory=dict)


def group_scored_chunks(store, row_ids, scores, threshold: float | None = None) -> list[list[ArticleGroup]]:
    """Group scored chunks into articles on the device.

    ``store``: a B200Store; ``row_ids`` int64 [Q, n] (global row ids, -1 = pad) and ``scores`` f32 [Q, n]
    in candidate order.  Returns, per query, the articles best-first with the reference's rank fields."""
    import torch

    ids = np.ascontiguousarray(row_ids, dtype=np.int64)
    sc = np.ascontiguousarray(scores, dtype=np.float32)
    if ids.ndim == 1:
        ids, sc = ids[None, :], sc[None, :]
    qn, n = ids.shape
    dense = store.dense
    if dense is None:
        return [[] for _ in range(qn)]
    dev = torch.device(f"cuda:{dense.device}")
    res = dense.multivector(torch.from_numpy(ids).to(dev).view(qn, 1, n), torch.from_numpy(sc).to(dev).view(qn, 1, n))
    torch.cuda.synchronize(dev)
    res = res.cpu()
    out: list[list[ArticleGroup]] = []
    for q in range(qn):
        g = int(res.grp_n[q])
        cand_grp = res.cand_grp[q].numpy()
        cand_ids = res.cand_ids[q].numpy()
        arts = []
        for gi in res.grp_order[q, :g].tolist():
            score = float(res.grp_max[q, gi])
            if threshold is not None and not (score >= threshold):
                continue
            rows = [int(r) for r in cand_ids[cand_grp == gi]]
            first = rows[0] - store._id_offset
            arts.append(ArticleGroup(kb_id=store.key_of_gid(int(res.grp_gid[q, gi])), score=score, rows=rows,
                                     metadata=dict(store._metas[first] or {})))
        for idx, a in enumerate(arts):
            a.article_rank = idx
            a.normalized_rank = idx / (len(arts) - 1) if len(arts) > 1 else 0.0
        out.append(arts)
    return out


# ------------------------------------------------------------------------------------------------
# confidence statistics over the scores of one query (retrieval/confidence.py:13-64 of the reference)
# ------------------------------------------------------------------------------------------------
def retrieval_confidence(scores, relevance_threshold: float | None = None, mean_top_k: int = 5) -> dict[str, Any]:
    """Same dict as the reference's ``compute_retrieval_confidence`` for one query's scores (``None`` entries are
    skipped): top_score, mean of the top ``mean_top_k``, gap between the top and the median score, how many reach
    the threshold (default 0.5), and the reference's conservative ``likely_relevant`` rule."""
    threshold = 0.5 if relevance_threshold is None else float(relevance_threshold)
    vals = np.asarray([float(s) for s in scores if s is not None], dtype=np.float64)
    if vals.size == 0:
        return {"top_score": 0.0, "mean_top_k": 0.0, "score_gap": 0.0, "n_above_threshold": 0, "likely_relevant": False}
    desc = np.sort(vals)[::-1]
    top = float(desc[0])
    head = desc[: max(1, int(mean_top_k))]
    mean_top = float(head.sum() / head.size)
    n = desc.size
    med = float(desc[n // 2]) if n % 2 else float((desc[n // 2 - 1] + desc[n // 2]) / 2.0)
    gap = top - med
    n_above = int((desc >= threshold).sum())
    return {"top_score": top, "mean_top_k": mean_top, "score_gap": gap, "n_above_threshold": n_above,
            "likely_relevant": bool(top >= threshold and (gap >= 0.05 or n_above >= 2))}


def retrieval_confidence_batch(scores, counts=None, relevance_threshold: float | None = None, mean_top_k: int = 5):
    """The same statistics for Q queries at once: ``scores`` [Q, n] (e.g. ``MultiVectorResult.grp_max`` or the
    rerank scores), ``counts`` [Q] = valid entries per row (default: all).  Returns one dict per query."""
    sc = np.asarray(scores, dtype=np.float64)
    if sc.ndim == 1:
        sc = sc[None, :]
    cnt = np.full(sc.shape[0], sc.shape[1]) if counts is None else np.asarray(counts).astype(int)
    return [retrieval_confidence(sc[q, : cnt[q]], relevance_threshold, mean_top_k) for q in range(sc.shape[0])]


def normalized_confidence_from_traces(query_traces) -> float | None:
    """``compute_normalized_confidence_from_traces`` (confidence.py:67-117): min-max normalise the traces'
    ``confidence.top_score`` values and average them; 0.5 each when they are all equal; None without any."""
    raw = []
    for trace in query_traces or []:
        conf = trace.get("confidence") if isinstance(trace, dict) else None
        if isinstance(conf, dict) and isinstance(conf.get("top_score"), (int, float)):
            raw.append(float(conf["top_score"]))
    if not raw:
        return None
    lo, hi = min(raw), max(raw)
    norm = [(s - lo) / (hi - lo) for s in raw] if hi > lo else [0.5] * len(raw)
    return sum(norm) / len(norm)


# ------------------------------------------------------------------------------------------------
# the same reduction one level up: several tool calls' article lists -> one list
# (accumulate_articles_from_tool_results, rag_engine/tools/utils.py:70-152 of the reference)
# ------------------------------------------------------------------------------------------------
def merge_tool_results(result_lists: Sequence[Sequence[tuple[Any, float | None, Any]]], device: int = 0):
    """Deduplicate articles across tool calls by kb_id, keeping the occurrence with the HIGHEST score (the first
    one among equals, like the reference's strict ``>``), articles without a kb_id kept as they are, the result
    ordered by score descending with ties in first-appearance order (Python's stable sort on the dict's insertion
    order).  ``result_lists``: per tool call, ``(kb_id, rerank_score or None, payload)`` in the call's order.

    The reduction runs on the device through K4 (``cmw_multivector`` with one "segment" per tool call: group by
    key, max score, first appearance, stable score-descending order) -- the kernel does not care where ids and
    scores come from.  Returns ``[(kb_id, score, payload), ...]``."""
    import torch

    from . import _native as N

    flat: list[tuple[Any, float, Any]] = []
    gids: list[int] = []
    key_gid: dict[Any, int] = {}
    for lst in result_lists:
        for kb_id, score, payload in lst:
            sc = -float("inf") if score is None else float(score)
            if not kb_id:
                gid = len(key_gid) + 1_000_000 + len(flat)  # unique: never merged (utils.py:118-123)
            else:
                gid = key_gid.setdefault(kb_id, len(key_gid))
            flat.append((kb_id, sc, payload))
            gids.append(gid)
    n = len(flat)
    if n == 0:
        return []
    if n > 2048:
        raise ValueError(f"merge_tool_results: {n} articles exceed the kernel's 2048 entries per reduction")
    # dense group numbers, then ONE K4 call: ids = positions (all distinct), kb table = group of each position
    remap = {g: i for i, g in enumerate(dict.fromkeys(gids))}
    dev = torch.device(f"cuda:{device}")
    kb = torch.tensor([remap[g] for g in gids], dtype=torch.int32, device=dev)
    ids = torch.arange(n, dtype=torch.int64, device=dev).view(1, 1, n)
    # -inf scores (articles without a rerank_score) must still group and sort last: K4 orders by score bits
    sc = torch.tensor([f[1] for f in flat], dtype=torch.float32, device=dev).view(1, 1, n)
    out = {name: torch.empty((1, n), dtype=dt, device=dev) for name, dt in
           (("grp_gid", torch.int32), ("grp_max", torch.float32), ("grp_cnt", torch.int32), ("grp_first", torch.int32),
            ("grp_order", torch.int32), ("cand_grp", torch.int32))}
    grp_n = torch.zeros((1,), dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    N.check(N.lib().cmw_multivector(kb.data_ptr(), n, 0, ids.data_ptr(), sc.data_ptr(), 1, 1, n, 0, 0, None, None, None,
                                    None, out["cand_grp"].data_ptr(), out["grp_gid"].data_ptr(), out["grp_max"].data_ptr(),
                                    out["grp_cnt"].data_ptr(), out["grp_first"].data_ptr(), out["grp_order"].data_ptr(),
                                    grp_n.data_ptr(), stream), "cmw_multivector")
    torch.cuda.synchronize(dev)
    g = int(grp_n[0])
    cand_grp = out["cand_grp"][0].cpu().numpy()
    grp_max = out["grp_max"][0].cpu().numpy()
    merged = []
    for gi in out["grp_order"][0, :g].cpu().tolist():
        members = np.flatnonzero(cand_grp == gi)
        best = next((int(m) for m in members if np.float32(flat[m][1]) == grp_max[gi]), int(members[0]))
        merged.append((flat[best][0], flat[best][1] if flat[best][1] > -float("inf") else None, flat[best][2]))
    return merged
