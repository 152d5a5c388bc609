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
