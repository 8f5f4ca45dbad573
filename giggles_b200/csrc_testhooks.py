"""Builds a ``gg_result`` from plain arrays through the C ABI's own D2H-free path — used by the CPU
tests of ``gg_result_call`` (GT / GQ / GL rules) when no device is present."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi


def make_result(columns) -> capi.Result:
    L = capi.lib()
    L.gg_result_from_host.argtypes = [C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_double)]
    L.gg_result_from_host.restype = C.c_void_p
    off = np.zeros(len(columns) + 1, np.uint64)
    np.cumsum([len(c) for c in columns], out=off[1:])
    vals = np.concatenate([np.asarray(c, np.float64) for c in columns]) if columns else np.zeros(0)
    h = L.gg_result_from_host(len(columns), capi.as_ptr(off, C.c_uint64), capi.as_ptr(vals, C.c_double))
    return capi.Result(C.c_void_p(h))
