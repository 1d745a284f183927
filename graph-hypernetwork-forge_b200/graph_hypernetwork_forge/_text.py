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


_pyhost = None


def _pyhost_lib():
    """lib/libghf_pyhost.so (csrc/pyhost.c): one C pass over the list, keyed on object identity.  None when the
    helper was not built - the numpy formulation below gives the same result, only slower."""
    global _pyhost
    if _pyhost is None:
        import ctypes
        import os
        path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "lib", "libghf_pyhost.so")
        try:
            lib = ctypes.PyDLL(path)
            lib.ghf_collapse_pylist.restype = ctypes.c_int64
            lib.ghf_collapse_pylist.argtypes = [ctypes.py_object, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
            lib.ghf_collapse_pylist_mt.restype = ctypes.c_int64
            lib.ghf_collapse_pylist_mt.argtypes = [ctypes.py_object, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                                   ctypes.c_int]
            _pyhost = lib
        except OSError:
            _pyhost = False
    return _pyhost or None


def collapse_by_identity(texts: List[str], out: np.ndarray = None, threads: int = 0):
    """-> (distinct objects in first-occurrence order, int32 map edge -> position in that list).
    `out`: int32 [len(texts)] buffer for the map (e.g. pinned memory, so that it goes to the device without another
    host copy); `threads`: worker threads of the C pass (0 = one per 1M entries, at most 16)."""
    n = len(texts)
    lib = _pyhost_lib() if type(texts) is list else None
    if out is not None and (out.dtype != np.int32 or out.shape != (n,) or not out.flags.c_contiguous):
        raise ValueError("collapse_by_identity: out must be a contiguous int32 array with one entry per text")
    if lib is not None:
        edge_map = np.empty(n, dtype=np.int32) if out is None else out
        first = np.empty(min(n, 1 << 22), dtype=np.int64)
        k = int(lib.ghf_collapse_pylist_mt(texts, edge_map.ctypes.data, first.ctypes.data, first.size, int(threads)))
        if k >= 0:
            return [texts[i] for i in first[:k]], edge_map
    ids = np.fromiter(map(id, texts), dtype=np.int64, count=n)
    _, first, inverse = np.unique(ids, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")          # distinct objects by first occurrence
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    objs = [texts[i] for i in first[order]]
    if out is not None:
        out[:] = rank[inverse]
        return objs, out
    return objs, rank[inverse].astype(np.int32)


def pack_texts(texts: List[str], edge_map_out: np.ndarray = None):
    """-> (utf8 bytes, offsets, edge_to_string or None).

    `edge_to_string[e]` indexes the packed strings; None means the identity map
    (every edge has its own packed string).  `edge_map_out`: buffer for the map (see `collapse_by_identity`).
    """
    if len(texts) < _IDENTITY_THRESHOLD:
        data, offsets = pack_utf8(texts)
        return data, offsets, None
    objs, edge_map = collapse_by_identity(texts, out=edge_map_out)
    data, offsets = pack_utf8(objs)
    return data, offsets, edge_map
