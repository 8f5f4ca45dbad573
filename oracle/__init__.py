"""TEST INFRASTRUCTURE — loaders for the CPU oracle (oracle/hmm_oracle.c, oracle/edit_oracle.c) and for the compiled
reference (oracle/_ref/libgiggles_ref.so, built by oracle/build_ref.sh from the reference's own
sources).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
may import this package; the product (giggles_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(_HERE, "_build", "libhmm_oracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libgiggles_ref.so")
PYREF = os.path.join(_HERE, "_ref", "pyref")


def build(with_python_ref: bool = False) -> None:
    subprocess.check_call(["make", "-s", "-C", _HERE, "_build/libhmm_oracle.so"])
    args = ["bash", os.path.join(_HERE, "build_ref.sh")] + (["--python"] if with_python_ref else [])
    subprocess.check_call(args)


def _ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _problem_args(p):
    return [p.n_columns, p.n_haps, _ptr(p.positions, C.c_uint32), _ptr(p.n_alleles, C.c_uint32),
            _ptr(p.allele_ref, C.c_int32), _ptr(p.recomb, C.c_float), p.n_reads, _ptr(p.read_tag, C.c_int8),
            _ptr(p.read_entry_offset, C.c_uint32), _ptr(p.entry_column, C.c_uint32), _ptr(p.entry_allele, C.c_int32),
            _ptr(p.entry_score_offset, C.c_uint64), _ptr(p.entry_scores, C.c_double), int(p.reg_const),
            C.c_double(p.base_const)]


_oracle = None


def oracle_lib():
    global _oracle
    if _oracle is None:
        if not os.path.exists(ORACLE_SO):
            build()
        _oracle = C.CDLL(ORACLE_SO)
        _oracle.oracle_hmm_run.restype = C.c_int
        _oracle.oracle_compatible_flat.restype = C.c_uint32
    return _oracle


def oracle_run(p) -> np.ndarray:
    """Genotype posteriors of problem ``p`` (flat, offsets ``p.gl_offsets()``) from the C restatement."""
    off = p.gl_offsets()
    out = np.full(int(off[-1]), np.nan)
    rc = oracle_lib().oracle_hmm_run(*_problem_args(p), _ptr(off, C.c_uint64), _ptr(out, C.c_double))
    if rc != 0:
        raise RuntimeError(f"oracle_hmm_run failed ({rc})")
    return out


def oracle_edit_distance(s, t, maxdiff: int = -1) -> int:
    """C restatement of the reference's edit_distance (oracle/edit_oracle.c, follows giggles/align.pyx:15-107)."""
    # align.pyx:30-31,42-45: lengths are taken from the str (characters), the comparison runs over the UTF-8 bytes:
    # for non-ASCII text the reference looks at the first len(s) BYTES only
    sb = s.encode()[:len(s)] if isinstance(s, str) else bytes(s)
    tb = t.encode()[:len(t)] if isinstance(t, str) else bytes(t)
    L = oracle_lib()
    L.oracle_edit_distance.restype = C.c_int
    return int(L.oracle_edit_distance(sb, len(sb), tb, len(tb), int(maxdiff)))


def reference_edit_distance():
    """The reference's own Cython edit_distance (oracle/_ref/pyref, built by build_ref.sh --python) or None."""
    import importlib
    import sys
    if not os.path.isdir(os.path.join(PYREF, "giggles")):
        return None
    sys.path.insert(0, PYREF)
    try:
        return importlib.import_module("giggles.align").edit_distance
    except Exception:
        return None
    finally:
        sys.path.remove(PYREF)


def oracle_entry_emission(scores, reg_const=10, base_const=float(np.e)) -> np.ndarray:
    s = np.ascontiguousarray(scores, dtype=np.float64)
    out = np.empty(s.size, dtype=np.longdouble)
    oracle_lib().oracle_entry_emission(_ptr(s, C.c_double), C.c_uint32(s.size), int(reg_const), C.c_double(base_const),
                                       out.ctypes.data_as(C.c_void_p))
    return out.astype(np.float64)


def have_reference() -> bool:
    return os.path.exists(REF_SO)


_ref = None


def ref_lib():
    global _ref
    if _ref is None:
        _ref = C.CDLL(REF_SO)
        _ref.ref_hmm_run.restype = C.c_int
        _ref.ref_structure.restype = C.c_int
    return _ref


def reference_run(p, return_seconds: bool = False):
    """The UNMODIFIED reference GenotypeHMM (compiled from /root/reference/src) on problem ``p``."""
    off = p.gl_offsets()
    out = np.full(int(off[-1]), np.nan)
    secs = C.c_double(0.0)
    err = C.create_string_buffer(512)
    rc = ref_lib().ref_hmm_run(*_problem_args(p), _ptr(off, C.c_uint64), _ptr(out, C.c_double), C.byref(secs), err,
                               C.c_size_t(len(err)))
    if rc != 0:
        raise RuntimeError(err.value.decode(errors="replace"))
    return (out, secs.value) if return_seconds else out


def reference_entry_emission(scores, reg_const=10, base_const=float(np.e)) -> np.ndarray:
    s = np.ascontiguousarray(scores, dtype=np.float64)
    out = np.empty_like(s)
    ref_lib().ref_entry_emission(_ptr(s, C.c_double), C.c_uint32(s.size), int(reg_const), C.c_double(base_const),
                                 _ptr(out, C.c_double))
    return out


def reference_structure(p, max_enum_bits: int = 8):
    """Active / untagged read ids per column and the compatible-bipartition lists, from the reference's own
    ColumnIterator / Column classes."""
    Cn = p.n_columns
    ao = np.zeros(Cn + 1, np.uint32)
    uo = np.zeros(Cn + 1, np.uint32)
    co = np.zeros(Cn + 1, np.uint64)
    needed = np.zeros(3, np.uint64)
    err = C.create_string_buffer(512)
    caps = [1, 1, 1]
    for _ in range(2):
        a = np.zeros(caps[0], np.uint32)
        u = np.zeros(caps[1], np.uint32)
        cm = np.zeros(caps[2], np.uint32)
        rc = ref_lib().ref_structure(
            Cn, _ptr(p.positions, C.c_uint32), p.n_reads, _ptr(p.read_tag, C.c_int8),
            _ptr(p.read_entry_offset, C.c_uint32), _ptr(p.entry_column, C.c_uint32), p.n_haps, max_enum_bits,
            _ptr(ao, C.c_uint32), _ptr(a, C.c_uint32), C.c_uint64(caps[0]),
            _ptr(uo, C.c_uint32), _ptr(u, C.c_uint32), C.c_uint64(caps[1]),
            _ptr(co, C.c_uint64), _ptr(cm, C.c_uint32), C.c_uint64(caps[2]),
            _ptr(needed, C.c_uint64), err, C.c_size_t(len(err)))
        if rc != 0:
            raise RuntimeError(err.value.decode(errors="replace"))
        if all(int(needed[i]) <= caps[i] for i in range(3)):
            break
        caps = [max(1, int(x)) for x in needed]
    active = [a[ao[c]:ao[c + 1]].tolist() for c in range(Cn)]
    untagged = [u[uo[c]:uo[c + 1]].tolist() for c in range(Cn)]
    compat = []
    for c in range(Cn):
        lists, i, end = [], int(co[c]), int(co[c + 1])
        while i < end:
            n = int(cm[i])
            lists.append(cm[i + 1:i + 1 + n].tolist())
            i += 1 + n
        compat.append(lists)
    return active, untagged, compat
