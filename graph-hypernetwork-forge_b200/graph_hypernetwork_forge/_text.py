"""Host-side packing of relation strings for the device dedup / text encoder.

The reference keys its dedup on whole Python strings (hypergnn.py:264-265).  At
the C boundary strings travel as UTF-8 bytes + int64 offsets; UTF-8 is injective
on str, so byte equality is string equality.

A Python list of E strings usually holds only a few distinct *objects* (callers
build it as ``[names[r] for r in rel]``).  `pack_texts` therefore first collapses
the list by object identity (one vectorised pass over ``id()``), packs only the
distinct objects in first-occurrence order, and returns the map from edges to
those objects; the device then dedups the packed strings by content.  Because
both steps keep first-occurrence order, composing them yields exactly
``list(dict.fromkeys(edge_texts))`` ids.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

_IDENTITY_THRESHOLD = 2048  # below this a direct pack is cheaper than the id() pass


def pack_utf8(texts: Sequence[str]) -> Tuple[np.ndarray, np.ndarray]:
    """-> (uint8 bytes, int64 offsets[len+1]); surrogates pass through unchanged."""
    n = len(texts)
    offsets = np.zeros(n + 1, dtype=np.int64)
    if n == 0:
        return np.zeros(0, dtype=np.uint8), offsets
    joined = "".join(texts)
    if joined.isascii():
        blob = joined.encode("ascii")
        np.cumsum(np.fromiter(map(len, texts), dtype=np.int64, count=n), out=offsets[1:])
    else:
        parts = [t.encode("utf-8", "surrogatepass") for t in texts]
        blob = b"".join(parts)
        np.cumsum(np.fromiter(map(len, parts), dtype=np.int64, count=n), out=offsets[1:])
    return np.frombuffer(blob, dtype=np.uint8), offsets


def collapse_by_identity(texts: List[str]):
    """-> (distinct objects in first-occurrence order, int32 map edge -> position in that list)."""
    n = len(texts)
    ids = np.fromiter(map(id, texts), dtype=np.int64, count=n)
    _, first, inverse = np.unique(ids, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")          # distinct objects by first occurrence
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    objs = [texts[i] for i in first[order]]
    return objs, rank[inverse].astype(np.int32)


def pack_texts(texts: List[str]):
    """-> (utf8 bytes, offsets, edge_to_string or None).

    `edge_to_string[e]` indexes the packed strings; None means the identity map
    (every edge has its own packed string).
    """
    if len(texts) < _IDENTITY_THRESHOLD:
        data, offsets = pack_utf8(texts)
        return data, offsets, None
    objs, edge_map = collapse_by_identity(texts)
    data, offsets = pack_utf8(objs)
    return data, offsets, edge_map
