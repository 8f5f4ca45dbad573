"""ctypes binding of the C ABI (include/giggles_b200.h) -> libgiggles_b200.so.

This is the reference-side stub a maintainer would write in Python; the drop-in
C++ class (integration/genotypehmm.cpp, bound by the reference's own giggles/core.pyx)
calls the same entry points.  There is
no fallback: if the shared library is missing, importing :func:`lib` raises.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

from .problem import Problem, as_ptr

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GG_LIB", os.path.join(_HERE, "libgiggles_b200.so"))   # GG_LIB: A/B runs of two builds

GG_OK, GG_ERR_INVALID, GG_ERR_NO_DEVICE, GG_ERR_CUDA, GG_ERR_NOMEM, GG_ERR_UNSUPPORTED = range(6)
GG_MAX_ALLELES = 16
GG_MAX_UNTAGGED = 24


class GGError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(message)
        self.status = status


class gg_problem(C.Structure):
    _fields_ = [
        ("n_columns", C.c_uint32),
        ("n_haps", C.c_uint32),
        ("positions", C.POINTER(C.c_uint32)),
        ("n_alleles", C.POINTER(C.c_uint32)),
        ("allele_ref", C.POINTER(C.c_int32)),
        ("recomb", C.POINTER(C.c_float)),
        ("n_reads", C.c_uint32),
        ("read_tag", C.POINTER(C.c_int8)),
        ("read_entry_offset", C.POINTER(C.c_uint32)),
        ("entry_column", C.POINTER(C.c_uint32)),
        ("entry_allele", C.POINTER(C.c_int32)),
        ("entry_em_offset", C.POINTER(C.c_uint64)),
        ("entry_em", C.POINTER(C.c_double)),
    ]


class gg_options(C.Structure):
    _fields_ = [
        ("device", C.c_int32),
        ("state_budget_bytes", C.c_uint64),
        ("precision", C.c_int32),
        ("kernel_variant", C.c_int32),
        ("flags", C.c_uint32),
    ]


GG_OPT_SHARED_ARENA = 1


class gg_readset_view(C.Structure):
    _fields_ = [
        ("n_reads", C.c_uint32),
        ("read_variant_offset", C.POINTER(C.c_uint64)),
        ("variant_position", C.POINTER(C.c_int32)),
        ("variant_quality", C.POINTER(C.c_int32)),
        ("read_source_id", C.POINTER(C.c_int32)),
    ]


class gg_stats(C.Structure):
    _fields_ = [
        ("n_columns", C.c_uint64),
        ("state_columns", C.c_uint64),
        ("algorithmic_bytes", C.c_uint64),
        ("scheduled_bytes", C.c_uint64),
        ("n_launches", C.c_uint64),
        ("n_backward_steps", C.c_uint64),
        ("n_forward_steps", C.c_uint64),
        ("state_bytes_reserved", C.c_uint64),
        ("max_untagged", C.c_uint32),
        ("checkpoint_levels", C.c_uint32),
        ("last_sweep_ms", C.c_float),
        ("last_fwdbwd_kernel_ms", C.c_float),
        ("last_fwd_ms", C.c_float),
        ("last_bwd_ms", C.c_float),
        ("last_emit_ms", C.c_float),
        ("create_plan_ms", C.c_float),
        ("create_total_ms", C.c_float),
        ("reserved", C.c_float),
    ]

    def as_dict(self) -> dict:
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}


# every symbol include/giggles_b200.h declares (tests check that the library exports them all)
SYMBOLS = [
    "gg_hmm_run", "gg_hmm_run_batch", "gg_hmm_create", "gg_hmm_sweep", "gg_hmm_fetch", "gg_hmm_stats", "gg_hmm_destroy", "gg_arena_release", "gg_hmm_debug_scales", "gg_hmm_debug_op_times",
    "gg_result_likelihoods", "gg_result_n_columns", "gg_result_flat", "gg_result_free", "gg_result_reran_fp64", "gg_fp64_rerun_count", "gg_result_call", "gg_result_from_host",
    "gg_plan_build", "gg_plan_n_columns", "gg_plan_active", "gg_plan_untagged", "gg_plan_shared_mask_prev",
    "gg_plan_compatible_prev", "gg_plan_free",
    "gg_entry_emission", "gg_entry_emission_batch", "gg_transition_scalars", "gg_binomial_coefficient",
    "gg_schedule_dry_run", "gg_default_state_budget", "gg_version", "gg_device_count", "gg_device_copy_gbs",
    "gg_readselection", "gg_pyset_replay", "gg_edit_distance_batch",
]

_lib = None


def lib() -> C.CDLL:
    """Load libgiggles_b200.so (built by ``__graft_entry__.build()`` / ``make -C giggles_b200/csrc``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `make -C giggles_b200/csrc` "
            "(there is no CPU or PyTorch fallback for the GenotypeHMM path)")
    L = C.CDLL(LIB_PATH)
    vp, u32, u64, i32, dbl = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int32, C.c_double
    errargs = [C.c_char_p, C.c_size_t]
    L.gg_hmm_run.argtypes = [C.POINTER(gg_problem), C.POINTER(gg_options), C.POINTER(vp)] + errargs
    L.gg_hmm_run_batch.argtypes = [C.POINTER(gg_problem), u32, C.POINTER(gg_options), C.POINTER(vp)] + errargs
    L.gg_hmm_create.argtypes = [C.POINTER(gg_problem), C.POINTER(gg_options), C.POINTER(vp)] + errargs
    L.gg_hmm_sweep.argtypes = [vp, C.c_int] + errargs
    L.gg_hmm_fetch.argtypes = [vp, C.POINTER(vp)] + errargs
    L.gg_hmm_stats.argtypes = [vp, C.POINTER(gg_stats)]
    if hasattr(L, "gg_hmm_debug_scales"):
        L.gg_hmm_debug_scales.argtypes = [vp, C.POINTER(dbl), C.POINTER(dbl)]
    L.gg_hmm_destroy.argtypes = [vp]
    L.gg_hmm_destroy.restype = None
    L.gg_result_likelihoods.argtypes = [vp, u32, C.POINTER(u32)]
    L.gg_result_likelihoods.restype = C.POINTER(dbl)
    L.gg_result_n_columns.argtypes = [vp]
    L.gg_result_n_columns.restype = u32
    L.gg_result_flat.argtypes = [vp, C.POINTER(C.POINTER(u64))]
    L.gg_result_flat.restype = C.POINTER(dbl)
    L.gg_result_free.argtypes = [vp]
    L.gg_result_free.restype = None
    L.gg_result_reran_fp64.argtypes = [vp]
    L.gg_fp64_rerun_count.restype = u64
    L.gg_result_call.argtypes = [vp, dbl, C.POINTER(i32), C.POINTER(i32), C.POINTER(dbl)]
    L.gg_plan_build.argtypes = [C.POINTER(gg_problem), C.POINTER(vp)] + errargs
    L.gg_plan_n_columns.argtypes = [vp]
    L.gg_plan_n_columns.restype = u32
    L.gg_plan_active.argtypes = [vp, u32, C.POINTER(u32)]
    L.gg_plan_active.restype = C.POINTER(u32)
    L.gg_plan_untagged.argtypes = [vp, u32, C.POINTER(u32)]
    L.gg_plan_untagged.restype = C.POINTER(u32)
    L.gg_plan_shared_mask_prev.argtypes = [vp, u32, C.POINTER(u32)]
    L.gg_plan_shared_mask_prev.restype = u32
    L.gg_plan_compatible_prev.argtypes = [vp, u32, u32, C.POINTER(u32), u32]
    L.gg_plan_compatible_prev.restype = u32
    L.gg_plan_free.argtypes = [vp]
    L.gg_plan_free.restype = None
    L.gg_readselection.argtypes = [C.POINTER(gg_readset_view), u32, C.POINTER(i32), u32, C.c_int, C.POINTER(u32),
                                   C.POINTER(u32)] + errargs
    L.gg_edit_distance_batch.argtypes = [C.POINTER(C.c_uint8), u64, C.POINTER(u64), C.POINTER(u32), C.POINTER(u64), C.POINTER(u32),
                                         C.POINTER(i32), u32, C.POINTER(i32), i32] + errargs
    L.gg_pyset_replay.argtypes = [C.POINTER(C.c_int64), C.c_int64, C.POINTER(C.c_int64), C.c_int64]
    L.gg_pyset_replay.restype = C.c_int64
    L.gg_entry_emission.argtypes = [C.POINTER(dbl), u32, i32, dbl, C.POINTER(dbl)]
    L.gg_entry_emission.restype = None
    L.gg_entry_emission_batch.argtypes = [C.POINTER(dbl), C.POINTER(u64), u64, i32, dbl, C.POINTER(dbl)]
    L.gg_entry_emission_batch.restype = None
    L.gg_transition_scalars.argtypes = [C.c_float, u32, C.POINTER(dbl), C.POINTER(dbl)]
    L.gg_transition_scalars.restype = None
    L.gg_binomial_coefficient.argtypes = [i32, i32]
    L.gg_binomial_coefficient.restype = i32
    L.gg_schedule_dry_run.argtypes = [u32, C.POINTER(u64), C.POINTER(u64), u64, C.POINTER(u64), C.POINTER(u32),
                                      C.POINTER(u64), C.POINTER(u64), C.POINTER(u64)] + errargs
    L.gg_default_state_budget.argtypes = [C.c_int]
    L.gg_default_state_budget.restype = u64
    L.gg_device_copy_gbs.argtypes = [C.c_int, u64, C.c_int]
    L.gg_device_copy_gbs.restype = dbl
    L.gg_version.restype = C.c_char_p
    L.gg_device_count.restype = C.c_int
    _lib = L
    return L


def _check(rc: int, err) -> None:
    if rc != GG_OK:
        raise GGError(rc, err.value.decode(errors="replace") or f"gg error {rc}")


def entry_emission(scores, reg_const: int = 10, base_const: float = float(np.e)) -> np.ndarray:
    """Entry::set_emission_score (reference src/entry.cpp:50-64)."""
    s = np.ascontiguousarray(scores, dtype=np.float64)
    out = np.empty_like(s)
    lib().gg_entry_emission(as_ptr(s, C.c_double), s.size, int(reg_const), float(base_const), as_ptr(out, C.c_double))
    return out


def transition_scalars(recombcost: float, n_haps: int):
    pr, qr = C.c_double(), C.c_double()
    lib().gg_transition_scalars(C.c_float(recombcost), n_haps, C.byref(pr), C.byref(qr))
    return pr.value, qr.value


class CProblem:
    """A :class:`Problem` laid out as ``gg_problem``: raw scores are turned into the normalised
    emission probabilities of ``Entry::set_emission_score`` by ``gg_entry_emission``."""

    def __init__(self, p: Problem):
        self.p = p
        L = lib()
        em = np.empty_like(p.entry_scores)
        L.gg_entry_emission_batch(as_ptr(p.entry_scores, C.c_double), as_ptr(p.entry_score_offset, C.c_uint64),
                                  p.n_entries, p.reg_const, p.base_const, as_ptr(em, C.c_double))
        self.entry_em = em
        self.struct = gg_problem(
            p.n_columns, p.n_haps,
            as_ptr(p.positions, C.c_uint32), as_ptr(p.n_alleles, C.c_uint32), as_ptr(p.allele_ref, C.c_int32),
            as_ptr(p.recomb, C.c_float), p.n_reads, as_ptr(p.read_tag, C.c_int8),
            as_ptr(p.read_entry_offset, C.c_uint32), as_ptr(p.entry_column, C.c_uint32),
            as_ptr(p.entry_allele, C.c_int32), as_ptr(p.entry_score_offset, C.c_uint64), as_ptr(em, C.c_double))


@dataclass
class Calls:
    gt_index: np.ndarray   # int32 [C], -1 = no call
    gq: np.ndarray         # int32 [C], -1 = undefined
    gl_log10: np.ndarray   # float64 flat


class Result:
    def __init__(self, handle):
        self._h = handle
        L = lib()
        offs = C.POINTER(C.c_uint64)()
        flat = L.gg_result_flat(handle, C.byref(offs))
        n = L.gg_result_n_columns(handle)
        self.offsets = np.ctypeslib.as_array(offs, shape=(n + 1,)).copy() if n else np.zeros(1, np.uint64)
        total = int(self.offsets[-1])
        self.values = np.ctypeslib.as_array(flat, shape=(total,)).copy() if total else np.zeros(0)
        self.reran_fp64 = bool(L.gg_result_reran_fp64(handle))

    def column(self, c: int) -> np.ndarray:
        return self.values[int(self.offsets[c]):int(self.offsets[c + 1])]

    def call(self, gt_qual_threshold: float = 0.0) -> Calls:
        n = len(self.offsets) - 1
        gt = np.empty(n, np.int32)
        gq = np.empty(n, np.int32)
        gl = np.empty(self.values.shape[0], np.float64)
        lib().gg_result_call(self._h, float(gt_qual_threshold), as_ptr(gt, C.c_int32), as_ptr(gq, C.c_int32),
                             as_ptr(gl, C.c_double))
        return Calls(gt, gq, gl)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().gg_result_free(self._h)
            self._h = None


def _options(device=-1, budget=0, precision=0, kernel_variant=0, flags=0):
    return gg_options(int(device), int(budget), int(precision), int(kernel_variant), int(flags))


def hmm_run(p: Problem | CProblem, device: int = -1, budget: int = 0, precision: int = 0,
            kernel_variant: int = 0, shared_arena: bool = False) -> Result:
    """``gg_hmm_run``: host buffers in, host posteriors out (what ``new GenotypeHMM(...)`` does in the reference)."""
    cp = p if isinstance(p, CProblem) else CProblem(p)
    err = C.create_string_buffer(512)
    out = C.c_void_p()
    o = _options(device, budget, precision, kernel_variant, GG_OPT_SHARED_ARENA if shared_arena else 0)
    _check(lib().gg_hmm_run(C.byref(cp.struct), C.byref(o), C.byref(out), err, len(err)), err)
    return Result(out)


def hmm_run_batch(problems, device: int = -1, budget: int = 0, precision: int = 0, kernel_variant: int = 0):
    """``gg_hmm_run_batch``: independent chains on one device — small chains side by side, HBM-filling ones one at a
    time with the planning of the next overlapped."""
    cps = [p if isinstance(p, CProblem) else CProblem(p) for p in problems]
    arr = (gg_problem * len(cps))(*[cp.struct for cp in cps])
    outs = (C.c_void_p * len(cps))()
    err = C.create_string_buffer(512)
    o = _options(device, budget, precision, kernel_variant)
    rc = lib().gg_hmm_run_batch(arr, len(cps), C.byref(o), outs, err, len(err))
    res = [Result(C.c_void_p(outs[i])) if outs[i] else None for i in range(len(cps))]
    _check(rc, err)
    return res


def hmm_run_many(problems, device: int = -1, workers: int = 4, precision: int = 0):
    """Independent chains run CONCURRENTLY on one device, one ``gg_hmm_run`` per host thread (ctypes releases the GIL;
    every call has its own stream) — the same thing :func:`hmm_run_batch` does inside the library with its own
    threads, for callers that already have a thread per chromosome.  Meant for small chains (all-haplotagged
    chromosomes are one persistent CTA each); chains whose columns fill HBM must not overlap and belong in
    :func:`hmm_run_batch`.  Results are identical to one call per chain, in the same order."""
    from concurrent.futures import ThreadPoolExecutor
    cps = [p if isinstance(p, CProblem) else CProblem(p) for p in problems]
    if not cps:
        return []
    with ThreadPoolExecutor(max_workers=max(1, min(workers, len(cps)))) as ex:
        return list(ex.map(lambda cp: hmm_run(cp, device=device, precision=precision), cps))


class Hmm:
    """Staged form: ``create`` (plan + H2D), ``sweep`` (device only), ``fetch`` (D2H)."""

    def __init__(self, p: Problem | CProblem, device: int = -1, budget: int = 0, precision: int = 0,
                 kernel_variant: int = 0, shared_arena: bool = False):
        self.cp = p if isinstance(p, CProblem) else CProblem(p)
        err = C.create_string_buffer(512)
        self._h = C.c_void_p()
        o = _options(device, budget, precision, kernel_variant, GG_OPT_SHARED_ARENA if shared_arena else 0)
        _check(lib().gg_hmm_create(C.byref(self.cp.struct), C.byref(o), C.byref(self._h), err, len(err)), err)

    def sweep(self, timed_kernels: bool = False) -> None:
        err = C.create_string_buffer(512)
        _check(lib().gg_hmm_sweep(self._h, int(timed_kernels), err, len(err)), err)

    def fetch(self) -> Result:
        err = C.create_string_buffer(512)
        out = C.c_void_p()
        _check(lib().gg_hmm_fetch(self._h, C.byref(out), err, len(err)), err)
        return Result(out)

    def debug_scales(self):
        n = self.cp.p.n_columns
        a, b = np.zeros(n), np.zeros(n)
        lib().gg_hmm_debug_scales(self._h, as_ptr(a, C.c_double), as_ptr(b, C.c_double))
        return a, b

    def debug_op_times(self, cap=1 << 20):
        k = np.zeros(cap, np.uint32); c = np.zeros(cap, np.uint32); ms = np.zeros(cap, np.float32)
        f = lib().gg_hmm_debug_op_times
        f.restype = C.c_uint64
        f.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_float), C.c_uint64]
        n = min(int(f(self._h, as_ptr(k, C.c_uint32), as_ptr(c, C.c_uint32), as_ptr(ms, C.c_float), cap)), cap)
        return k[:n], c[:n], ms[:n]

    def stats(self) -> dict:
        s = gg_stats()
        lib().gg_hmm_stats(self._h, C.byref(s))
        return s.as_dict()

    def close(self):
        if getattr(self, "_h", None):
            lib().gg_hmm_destroy(self._h)
            self._h = None

    __del__ = close


class PlanView:
    """Column structure as the product derives it (``gg_plan_*``), for the structure parity test."""

    def __init__(self, p: Problem | CProblem):
        self.cp = p if isinstance(p, CProblem) else CProblem(p)
        err = C.create_string_buffer(512)
        self._h = C.c_void_p()
        _check(lib().gg_plan_build(C.byref(self.cp.struct), C.byref(self._h), err, len(err)), err)
        self.n_columns = lib().gg_plan_n_columns(self._h)

    def _ids(self, fn, c):
        n = C.c_uint32()
        ptr = fn(self._h, c, C.byref(n))
        return [ptr[i] for i in range(n.value)]

    def active(self, c):
        return self._ids(lib().gg_plan_active, c)

    def untagged(self, c):
        return self._ids(lib().gg_plan_untagged, c)

    def shared_mask_prev(self, c):
        n = C.c_uint32()
        m = lib().gg_plan_shared_mask_prev(self._h, c, C.byref(n))
        return m, n.value

    def compatible_prev(self, c, b, cap=1 << 16):
        buf = (C.c_uint32 * cap)()
        n = lib().gg_plan_compatible_prev(self._h, c, b, buf, cap)
        return [buf[i] for i in range(min(n, cap))]

    def __del__(self):
        if getattr(self, "_h", None):
            lib().gg_plan_free(self._h)
            self._h = None


def readselection_flat(read_variant_offset, variant_position, variant_quality, read_source_id, max_cov: int,
                       preferred_source_ids=None, bridging: bool = True) -> np.ndarray:
    """``gg_readselection`` on flat arrays; returns the selected read indices, increasing."""
    off = np.ascontiguousarray(read_variant_offset, dtype=np.uint64)
    pos = np.ascontiguousarray(variant_position, dtype=np.int32)
    qual = np.ascontiguousarray(variant_quality, dtype=np.int32)
    src = np.ascontiguousarray(read_source_id, dtype=np.int32)
    n = off.size - 1
    if n < 0 or src.size != n or pos.size != qual.size or (n >= 0 and int(off[-1]) != pos.size):
        raise ValueError("readselection_flat: inconsistent array sizes")
    view = gg_readset_view(n, as_ptr(off, C.c_uint64), as_ptr(pos, C.c_int32), as_ptr(qual, C.c_int32),
                           as_ptr(src, C.c_int32))
    pref = None if preferred_source_ids is None else np.ascontiguousarray(sorted(preferred_source_ids), dtype=np.int32)
    out = np.empty(max(n, 1), dtype=np.uint32)
    cnt = C.c_uint32()
    err = C.create_string_buffer(512)
    rc = lib().gg_readselection(C.byref(view), int(max_cov), None if pref is None else as_ptr(pref, C.c_int32),
                                0 if pref is None else pref.size, int(bool(bridging)), as_ptr(out, C.c_uint32),
                                C.byref(cnt), err, len(err))
    if rc == 2:
        raise ValueError(err.value.decode())
    _check(rc, err)
    return out[:cnt.value].copy()


def pyset_replay(ops, cap: int) -> list:
    """Test hook: iteration orders of the order-exact Python-set model after a script of operations."""
    a = np.ascontiguousarray(ops, dtype=np.int64)
    out = np.empty(cap, dtype=np.int64)
    n = lib().gg_pyset_replay(as_ptr(a, C.c_int64), a.size, as_ptr(out, C.c_int64), cap)
    if n < 0:
        raise GGError(1, f"gg_pyset_replay: {n}")
    return out[:n].tolist()


def schedule_dry_run(col_bytes, fg_bytes, budget):
    cb = np.ascontiguousarray(col_bytes, dtype=np.uint64)
    fb = np.ascontiguousarray(fg_bytes, dtype=np.uint64)
    arena, levels, nb, nf, bb = C.c_uint64(), C.c_uint32(), C.c_uint64(), C.c_uint64(), C.c_uint64()
    err = C.create_string_buffer(512)
    rc = lib().gg_schedule_dry_run(cb.size, as_ptr(cb, C.c_uint64), as_ptr(fb, C.c_uint64), int(budget),
                                   C.byref(arena), C.byref(levels), C.byref(nb), C.byref(nf), C.byref(bb), err, len(err))
    _check(rc, err)
    return dict(arena_bytes=arena.value, levels=levels.value, n_bwd=nb.value, n_fwd=nf.value, bwd_col_bytes=bb.value)
