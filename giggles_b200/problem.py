"""Flat (structure-of-arrays) form of one GenotypeHMM call.

This is the host-side image of ``gg_problem`` (include/giggles_b200.h): one
chromosome's worth of input to ``GenotypeHMM::GenotypeHMM``
(reference src/genotypehmm.cpp:19-51, ctor at src/genotypehmm.h:142) after the
ReadSet has been sorted (reference src/readset.cpp:44-53).  ``entry_scores`` are
the RAW per-allele alignment scores handed to ``Read.add_variant``
(reference giggles/core.pyx:208-214); they become emission probabilities
through ``Entry::set_emission_score`` (reference src/entry.cpp:50-64) which the
product implements as ``gg_entry_emission``.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np


@dataclass
class Problem:
    positions: np.ndarray          # u32 [C], strictly increasing
    n_alleles: np.ndarray          # u32 [C]
    allele_ref: np.ndarray         # i32 [C, R], -1 = missing
    recomb: np.ndarray             # f32 [max(C-1, 0)]
    read_tag: np.ndarray           # i8  [N], -1 untagged / 0 = H1 / 1 = H2
    read_entry_offset: np.ndarray  # u32 [N+1]
    entry_column: np.ndarray       # u32 [E]
    entry_allele: np.ndarray       # i32 [E]
    entry_score_offset: np.ndarray # u64 [E+1]
    entry_scores: np.ndarray       # f64 [sum n_alleles of entries], raw scores
    reg_const: int = 10            # --em-reg-constant (reference giggles/cli/genotype.py:116)
    base_const: float = float(np.e)  # --em-score-to-prob-base-constant (:117)
    meta: dict = field(default_factory=dict)

    def __post_init__(self):
        self.positions = np.ascontiguousarray(self.positions, dtype=np.uint32)
        self.n_alleles = np.ascontiguousarray(self.n_alleles, dtype=np.uint32)
        self.allele_ref = np.ascontiguousarray(self.allele_ref, dtype=np.int32)
        if self.allele_ref.ndim != 2:
            raise ValueError("allele_ref must be [C, R]")
        self.recomb = np.ascontiguousarray(self.recomb, dtype=np.float32)
        self.read_tag = np.ascontiguousarray(self.read_tag, dtype=np.int8)
        self.read_entry_offset = np.ascontiguousarray(self.read_entry_offset, dtype=np.uint32)
        self.entry_column = np.ascontiguousarray(self.entry_column, dtype=np.uint32)
        self.entry_allele = np.ascontiguousarray(self.entry_allele, dtype=np.int32)
        self.entry_score_offset = np.ascontiguousarray(self.entry_score_offset, dtype=np.uint64)
        self.entry_scores = np.ascontiguousarray(self.entry_scores, dtype=np.float64)

    # ---- sizes -----------------------------------------------------------------
    @property
    def n_columns(self) -> int:
        return int(self.positions.shape[0])

    @property
    def n_haps(self) -> int:
        return int(self.allele_ref.shape[1])

    @property
    def n_reads(self) -> int:
        return int(self.read_tag.shape[0])

    @property
    def n_entries(self) -> int:
        return int(self.entry_column.shape[0])

    def gl_offsets(self) -> np.ndarray:
        """Flat offsets of the per-column genotype tables: n(n+1)/2 entries each
        (reference src/genotypehmm.cpp:36: binomial_coefficient(n+1, n-1))."""
        n = self.n_alleles.astype(np.uint64)
        off = np.zeros(self.n_columns + 1, dtype=np.uint64)
        np.cumsum(n * (n + 1) // 2, out=off[1:])
        return off

    def untagged_per_column(self) -> np.ndarray:
        """u_c = number of untagged reads spanning column c (reference src/column.cpp:13-17)."""
        u = np.zeros(self.n_columns + 1, dtype=np.int64)
        off = self.read_entry_offset
        for k in range(self.n_reads):
            if self.read_tag[k] != -1:
                continue
            first = int(self.entry_column[off[k]])
            last = int(self.entry_column[off[k + 1] - 1])
            u[first] += 1
            u[last + 1] -= 1
        return np.cumsum(u[:-1])

    def state_columns(self) -> int:
        """sum_c 2^{u_c} R^2 (reference src/column.cpp:51-53)."""
        u = self.untagged_per_column()
        return int(np.sum(np.left_shift(np.int64(1), u)) * self.n_haps * self.n_haps)

    # ---- (de)serialisation for tests/golden ----------------------------------
    _ARRAYS = ("positions", "n_alleles", "allele_ref", "recomb", "read_tag", "read_entry_offset",
               "entry_column", "entry_allele", "entry_score_offset", "entry_scores")

    def to_npz_dict(self) -> dict:
        d = {k: getattr(self, k) for k in self._ARRAYS}
        d["reg_const"] = np.int64(self.reg_const)
        d["base_const"] = np.float64(self.base_const)
        return d

    @classmethod
    def from_npz_dict(cls, d) -> "Problem":
        kw = {k: d[k] for k in cls._ARRAYS}
        return cls(reg_const=int(d["reg_const"]), base_const=float(d["base_const"]), **kw)


def as_ptr(a: np.ndarray, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))
