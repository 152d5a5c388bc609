"""Golden vectors for the rerank-side plumbing, produced by the REFERENCE's own functions:

* ``compute_retrieval_confidence`` / ``compute_normalized_confidence_from_traces``
  (rag_engine/retrieval/confidence.py:13-117),
* ``accumulate_articles_from_tool_results`` (rag_engine/tools/utils.py:70-152; the module is loaded by file path --
  its package ``__init__`` pulls in langchain, which is not installed; the function itself needs only ``Article``).

Build container only (needs /root/reference):  python tests/golden/make_golden_f3.py
"""
from __future__ import annotations

import importlib.util
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402


def main() -> None:
    mg._install_shim(tempfile.mkdtemp(prefix="golden_f3_"))
    from rag_engine.retrieval.confidence import (compute_normalized_confidence_from_traces,
                                                 compute_retrieval_confidence)

    spec = importlib.util.spec_from_file_location("ref_tools_utils", os.path.join(mg.REF, "rag_engine", "tools", "utils.py"))
    utils = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(utils)

    rng = np.random.default_rng(20261018)
    conf = []
    for n in (0, 1, 2, 3, 5, 6, 10, 20, 60):
        for thr in (None, 0.3, 0.5, 0.9):
            for mk in (1, 5):
                scores = [None if rng.random() < 0.1 else float(np.round(rng.uniform(-0.2, 1.3), 4)) for _ in range(n)]
                if n >= 3 and rng.random() < 0.3:
                    scores[1] = scores[0]  # ties
                out = compute_retrieval_confidence([(None, s) for s in scores], relevance_threshold=thr, mean_top_k=mk)
                conf.append({"scores": scores, "threshold": thr, "mean_top_k": mk, "expected": out})
    traces = []
    for case in ([], [{"confidence": {"top_score": 0.8}}, {"confidence": {"top_score": 1.0}}, {"confidence": {"top_score": 1.2}}],
                 [{"confidence": {"top_score": 0.7}}, {"confidence": {"top_score": 0.7}}],
                 [{"confidence": None}, {"x": 1}, {"confidence": {"top_score": "bad"}}],
                 [{"confidence": {"top_score": 2}}, {"confidence": {}}, {"confidence": {"top_score": -1.5}}]):
        traces.append({"traces": case, "expected": compute_normalized_confidence_from_traces(case)})

    merges = []
    for case_no in range(24):
        n_tools = int(rng.integers(1, 5))
        tools = []
        serial = 0
        for t in range(n_tools):
            arts = []
            for _ in range(int(rng.integers(0, 9))):
                kb = "" if rng.random() < 0.08 else str(int(rng.integers(100, 112)))
                meta = {}
                if rng.random() > 0.1:
                    # a coarse grid makes equal scores (ties) common
                    meta["rerank_score"] = float(np.round(rng.uniform(0.0, 1.2), 1))
                arts.append({"kb_id": kb, "title": f"t{serial}", "url": "u", "content": f"content-{serial}", "metadata": meta})
                serial += 1
            tools.append(json.dumps({"articles": arts}))
        got = utils.accumulate_articles_from_tool_results(tools)
        merges.append({"tool_results": tools,
                       "expected": [[a.kb_id, a.content, (a.metadata or {}).get("rerank_score")] for a in got]})
    out = {"about": "golden vectors from the reference's confidence.py and tools/utils.py (see make_golden_f3.py)",
           "confidence": conf, "normalized": traces, "tool_merge": merges}
    path = os.path.join(HERE, "f3_golden.json")
    with open(path, "w") as f:
        json.dump(out, f)
    print(path, len(conf), len(traces), len(merges))


if __name__ == "__main__":
    main()
