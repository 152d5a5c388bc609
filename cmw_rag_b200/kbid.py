"""kbId normalisation on the host -- the grouping key of the multi-vector reduction.

Mirrors ``extract_numeric_kbid`` (rag_engine/utils/metadata_utils.py:20-32 of the reference: the
leading run of decimal digits of ``str(kb_id)``, else None) and the key the retriever builds from
it (rag_engine/retrieval/retriever.py:236-239: falsy kbId -> chunk skipped; otherwise the numeric
prefix, falling back to ``str(kbId)``).  Strings stay on the host; the device sees a dense int32
group number per row (``kb_gid``), assigned at ingest.
"""
from __future__ import annotations


def extract_numeric_kbid(kb_id) -> str | None:
    if kb_id is None:
        return None
    s = str(kb_id)
    n = 0
    while n < len(s) and s[n].isdecimal():  # regex \d of the reference = Unicode decimal digits
        n += 1
    return s[:n] if n else None


def group_key(raw_kb_id) -> str | None:
    """None = "skip this chunk" (falsy kbId), else the normalised article key."""
    if not raw_kb_id:
        return None
    return extract_numeric_kbid(raw_kb_id) or str(raw_kb_id)
