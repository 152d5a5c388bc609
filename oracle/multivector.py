"""Multi-vector union / cap / kbId grouping -- CPU restatement (TEST INFRASTRUCTURE ONLY).

Follows ``rag_engine/retrieval/retriever.py`` of the reference:

* ``:185-194``  ordered union: segments in order, hits in rank order, first-seen wins,
  key = ``metadata["stable_id"]`` (one per corpus row, so row id here);
* ``:208-210``  pre-rerank cap ``candidates[:prl]`` when ``prl > 0``;
* ``:229-231``  no-reranker truncation ``scored_candidates[:qk]`` (``limit`` below);
* ``:234-242``  group by ``extract_numeric_kbid(kbId) or str(kbId)``, skip falsy kbId,
  keep members in order and the max score; dict order = first appearance;
* ``:307-316``  stable sort by score descending; ``normalized_rank = idx / (n - 1)``.

and ``rag_engine/utils/metadata_utils.py:20-32`` for the group key.

Pure-Python loops on purpose (small cases); written from the reference's *behaviour*,
checked against the reference's own class by ``tests/golden/make_golden.py``.
"""
from __future__ import annotations

import re
from typing import Sequence

import numpy as np

_LEADING_DIGITS = re.compile(r"^(\d+)")


def extract_numeric_kbid(kb_id) -> str | None:
    """Leading-digit run of ``str(kb_id)`` or None (metadata_utils.py:20-32)."""
    if kb_id is None:
        return None
    m = _LEADING_DIGITS.match(str(kb_id))
    return m.group(1) if m else None


def group_key(raw_kb_id) -> str | None:
    """Group key used at retriever.py:236-239; None means "skip this chunk"."""
    if not raw_kb_id:
        return None
    return extract_numeric_kbid(raw_kb_id) or str(raw_kb_id)


def union_dedup_cap(seg_ids: Sequence[Sequence[int]], seg_scores=None, prl: int = 0):
    """retriever.py:185-194 + 208-210 for one long query.

    ``seg_ids[s][r]`` is the row id at rank r of segment s (negative = padding, skipped).
    Returns (cand_ids, cand_first_scores, cand_best_scores): first-seen order; the score
    kept is the one of the first occurrence (the doc object the reference keeps);
    ``best`` is the maximum over all occurrences before the cap (an extra the GPU path
    exposes; it does not influence order).
    """
    seen: dict[int, int] = {}
    ids: list[int] = []
    first: list[float] = []
    best: list[float] = []
    for s, hits in enumerate(seg_ids):
        for r, rid in enumerate(hits):
            rid = int(rid)
            if rid < 0:
                continue
            sc = float(seg_scores[s][r]) if seg_scores is not None else 0.0
            pos = seen.get(rid)
            if pos is not None:
                if sc > best[pos]:
                    best[pos] = sc
                continue
            seen[rid] = len(ids)
            ids.append(rid)
            first.append(sc)
            best.append(sc)
    if prl and prl > 0 and len(ids) > prl:
        ids, first, best = ids[:prl], first[:prl], best[:prl]
    return ids, first, best


def group_by_kbid(cand_ids: Sequence[int], cand_scores: Sequence[float], kb_gid, limit: int = 0):
    """retriever.py:229-242 + 307 for one long query.

    ``kb_gid[row]`` is the dense group number of the row's normalised kbId, or a
    negative number for a falsy kbId (skipped, retriever.py:238).
    Returns a dict with groups in first-appearance order and the stable score-desc order.
    """
    if limit and limit > 0:
        cand_ids = cand_ids[:limit]
        cand_scores = cand_scores[:limit]
    index: dict[int, int] = {}
    gids: list[int] = []
    gmax: list[float] = []
    gcnt: list[int] = []
    gfirst: list[int] = []
    members: list[list[int]] = []
    cand_grp: list[int] = []
    for pos, (rid, sc) in enumerate(zip(cand_ids, cand_scores)):
        g = int(kb_gid[int(rid)])
        if g < 0:
            cand_grp.append(-1)
            continue
        gi = index.get(g)
        if gi is None:
            gi = len(gids)
            index[g] = gi
            gids.append(g)
            gmax.append(-float("inf"))
            gcnt.append(0)
            gfirst.append(pos)
            members.append([])
        members[gi].append(pos)
        gcnt[gi] += 1
        gmax[gi] = max(gmax[gi], float(sc))
        cand_grp.append(gi)
    # list.sort(key=score, reverse=True) is stable: ties keep first-appearance order
    order = sorted(range(len(gids)), key=lambda gi: gmax[gi], reverse=True)
    return {
        "gid": gids,
        "max": gmax,
        "cnt": gcnt,
        "first": gfirst,
        "members": members,
        "cand_grp": cand_grp,
        "order": order,
    }


def normalized_ranks(n: int) -> list[float]:
    """retriever.py:309-316."""
    if n <= 0:
        return []
    if n == 1:
        return [0.0]
    return [i / (n - 1) for i in range(n)]


def multivector_reduce(ids, scores, kb_gid, prl: int = 0, limit: int = 0):
    """Array-shaped restatement matching the CUDA kernel's outputs.

    ids int64[Q, S, k], scores f32[Q, S, k] -> dict of arrays, P = prl if 0 < prl < S*k else S*k:
      cand_ids i64[Q,P] (-1 pad), cand_scores f32[Q,P] (first seen), cand_best f32[Q,P],
      cand_n i32[Q], cand_grp i32[Q,P] (-1 = skipped / beyond limit / pad),
      grp_gid i32[Q,P], grp_max f32[Q,P], grp_cnt i32[Q,P], grp_first i32[Q,P],
      grp_order i32[Q,P], grp_n i32[Q]
    """
    ids = np.asarray(ids)
    scores = np.asarray(scores)
    qn, s, k = ids.shape
    p = prl if (prl and 0 < prl < s * k) else s * k
    out = {
        "cand_ids": np.full((qn, p), -1, np.int64),
        "cand_scores": np.full((qn, p), -np.inf, np.float32),
        "cand_best": np.full((qn, p), -np.inf, np.float32),
        "cand_n": np.zeros(qn, np.int32),
        "cand_grp": np.full((qn, p), -1, np.int32),
        "grp_gid": np.full((qn, p), -1, np.int32),
        "grp_max": np.full((qn, p), -np.inf, np.float32),
        "grp_cnt": np.zeros((qn, p), np.int32),
        "grp_first": np.full((qn, p), -1, np.int32),
        "grp_order": np.full((qn, p), -1, np.int32),
        "grp_n": np.zeros(qn, np.int32),
    }
    for q in range(qn):
        cid, cfirst, cbest = union_dedup_cap(ids[q], scores[q], prl)
        n = len(cid)
        out["cand_n"][q] = n
        out["cand_ids"][q, :n] = cid
        out["cand_scores"][q, :n] = np.asarray(cfirst, np.float32)
        out["cand_best"][q, :n] = np.asarray(cbest, np.float32)
        g = group_by_kbid(cid, [np.float32(x) for x in cfirst], kb_gid, limit)
        m = len(g["gid"])
        out["grp_n"][q] = m
        out["grp_gid"][q, :m] = g["gid"]
        out["grp_max"][q, :m] = np.asarray(g["max"], np.float32)
        out["grp_cnt"][q, :m] = g["cnt"]
        out["grp_first"][q, :m] = g["first"]
        out["grp_order"][q, :m] = g["order"]
        out["cand_grp"][q, : len(g["cand_grp"])] = g["cand_grp"]
    return out
