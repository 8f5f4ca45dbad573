"""Host-side mirror of the reference's ``giggles.align`` for the realignment step.

``edit_distance(s, t, maxdiff=-1)`` has the reference's signature and result (giggles/align.pyx:15-107) but runs on
the device; ``edit_distance_batch`` is the form the allele detection wants — the reference calls the aligner once per
read x variant x allele (giggles/variants.py:229-239), which is one batch per read or per chromosome here.
There is no CPU fallback: without the CUDA library / a device the call raises.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .problem import as_ptr


def _bytes(x) -> bytes:
    # the reference takes the length from the str and compares UTF-8 bytes (align.pyx:30-31,42-45): non-ASCII text
    # is cut to its first len(x) bytes there, hence here
    return x.encode()[:len(x)] if isinstance(x, str) else bytes(x)


def pack_pairs(pairs):
    """Flatten (s, t) pairs into the C ABI's image: one byte pool (every distinct string once — a read's query is
    compared against every allele) and per-pair offsets / lengths."""
    pairs = list(pairs)
    n = len(pairs)
    index: dict[bytes, tuple[int, int]] = {}
    chunks, pos = [], 0
    s_off = np.empty(n, np.uint64); s_len = np.empty(n, np.uint32)
    t_off = np.empty(n, np.uint64); t_len = np.empty(n, np.uint32)
    for i, (s, t) in enumerate(pairs):
        for arr_off, arr_len, x in ((s_off, s_len, _bytes(s)), (t_off, t_len, _bytes(t))):
            if x not in index:
                index[x] = (pos, len(x))
                chunks.append(x)
                pos += len(x)
            arr_off[i], arr_len[i] = index[x]
    pool = np.frombuffer(b"".join(chunks) or b"\0", dtype=np.uint8)
    return pool, pos, s_off, s_len, t_off, t_len


def edit_distance_packed(packed, maxdiff=-1, device: int = -1) -> np.ndarray:
    """One ``gg_edit_distance_batch`` call on the output of :func:`pack_pairs` (H2D / D2H inside the call)."""
    pool, pos, s_off, s_len, t_off, t_len = packed
    n = len(s_off)
    out = np.zeros(n, dtype=np.int32)
    if n == 0:
        return out
    md = np.full(n, int(maxdiff), np.int32) if np.isscalar(maxdiff) else np.ascontiguousarray(maxdiff, dtype=np.int32)
    if md.shape != (n,):
        raise ValueError("maxdiff: one value or one per pair")
    err = C.create_string_buffer(512)
    rc = capi.lib().gg_edit_distance_batch(as_ptr(pool, C.c_uint8), pos, as_ptr(s_off, C.c_uint64), as_ptr(s_len, C.c_uint32),
                                           as_ptr(t_off, C.c_uint64), as_ptr(t_len, C.c_uint32), as_ptr(md, C.c_int32), n,
                                           as_ptr(out, C.c_int32), int(device), err, len(err))
    capi._check(rc, err)
    return out


def edit_distance_batch(pairs, maxdiff=-1, device: int = -1) -> np.ndarray:
    """``pairs``: sequence of (s, t) (str or bytes).  ``maxdiff``: one int for all pairs or one per pair.
    Returns int32 distances, pair by pair, exactly what ``giggles.align.edit_distance`` returns for each."""
    return edit_distance_packed(pack_pairs(pairs), maxdiff, device)


def edit_distance(s, t, maxdiff: int = -1) -> int:
    """Drop-in for ``giggles.align.edit_distance`` (one pair; batch the calls where you can)."""
    return int(edit_distance_batch([(s, t)], maxdiff)[0])
