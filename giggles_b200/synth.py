"""Seeded synthetic GenotypeHMM problems.

The reference ships no data and its GAF/GFA/VCF front end cannot run without
pysam/pywfa, so workloads are generated at the level the hot path consumes:
what ``giggles genotype`` hands to ``GenotypeHMM`` after parsing, realignment,
read selection and ``ReadSet.sort()`` (reference giggles/cli/genotype.py:205-322).

* positions: exponential gaps, mean = region / C
* panel: Li-Stephens-like copying over a few founder haplotypes, a small
  fraction of missing (-1) entries, never a whole column
* recombination cost: reference giggles/pedigree.py:26
* sample: mosaic of two panel haplotypes; reads drawn from one of them
* reads: uniform starts, log-normal lengths, interior gaps, per-entry raw scores
  ``-edit distance``; coverage capped greedily at ``max_coverage`` so that the
  number of untagged active reads u_c never exceeds it (the role
  ``readselection`` plays in the reference, giggles/readselect.pyx:239-271)
"""
from __future__ import annotations

import numpy as np

from .problem import Problem


def uniform_recombination_cost(positions, recombrate=1.26, eff_pop_size=10):
    """reference giggles/pedigree.py:19-26 (python floats, then narrowed to fp32 by the ctor)."""
    pos = [int(x) for x in positions]
    return np.array(
        [(pos[i] - pos[i - 1]) * recombrate * eff_pop_size * (4 / (pow(10, 6))) for i in range(1, len(pos))],
        dtype=np.float64,
    ).astype(np.float32)


def make_problem(
    n_columns: int,
    n_haps: int,
    *,
    region_bp: int | None = None,
    coverage: float = 10.0,
    max_coverage: int = 15,
    read_len_mean: float = 15000.0,
    read_len_sigma: float = 0.2,
    accuracy: float = 0.99,
    tags: str = "none",            # "none" | "all" | "mixed"
    multiallelic_frac: float = 0.1,
    max_alleles: int = 8,
    min_alleles_multi: int = 3,
    missing_frac: float = 0.005,
    gap_frac: float = 0.02,
    unknown_allele_frac: float = 0.0,
    n_founders: int = 12,
    founder_switch: float = 0.02,
    window: bool = False,          # True: the region is a window inside a chromosome (reads may start before it)
    seed: int = 0,
    reg_const: int = 10,
    base_const: float = float(np.e),
) -> Problem:
    rng = np.random.default_rng(seed)
    C, R = int(n_columns), int(n_haps)
    if region_bp is None:
        region_bp = 200 * max(C, 1)
    # --- positions -------------------------------------------------------------
    gaps = np.maximum(1, np.round(rng.exponential(region_bp / max(C, 1), size=C))).astype(np.int64)
    positions = np.cumsum(gaps).astype(np.uint32)
    # --- alleles per column ----------------------------------------------------
    n_alleles = np.full(C, 2, dtype=np.uint32)
    multi = rng.random(C) < multiallelic_frac
    if max_alleles >= min_alleles_multi:
        n_alleles[multi] = rng.integers(min_alleles_multi, max_alleles + 1, size=int(multi.sum()))
    # --- panel: founders then copying ------------------------------------------
    nf = max(2, min(n_founders, max(R, 2)))
    founders = np.zeros((nf, C), dtype=np.int32)
    for c in range(C):
        n = int(n_alleles[c])
        # skewed allele frequencies: allele 0 common
        p = np.array([0.6] + [0.4 / (n - 1)] * (n - 1))
        founders[:, c] = rng.choice(n, size=nf, p=p)
    panel = np.zeros((C, R), dtype=np.int32)
    for h in range(R):
        f = int(rng.integers(nf))
        sw = rng.random(C) < founder_switch
        src = np.empty(C, dtype=np.int64)
        for c in range(C):
            if sw[c]:
                f = int(rng.integers(nf))
            src[c] = f
        panel[:, h] = founders[src, np.arange(C)]
    miss = rng.random((C, R)) < missing_frac
    for c in range(C):
        if miss[c].all():
            miss[c, int(rng.integers(R))] = False
    panel[miss] = -1
    # --- sample = mosaic of two panel haplotypes ---------------------------------
    truth = np.zeros((2, C), dtype=np.int32)
    for side in range(2):
        h = int(rng.integers(R))
        for c in range(C):
            if rng.random() < 0.01:
                h = int(rng.integers(R))
            a = panel[c, h]
            truth[side, c] = a if a >= 0 else 0
    # --- reads ---------------------------------------------------------------
    span_bp = int(positions[-1]) if C else 0
    n_cand = int(np.ceil(coverage * max(span_bp, 1) / read_len_mean * 1.3)) if C else 0
    lo = -int(2 * read_len_mean) if window else 0
    n_cand = int(n_cand * (span_bp - lo) / max(span_bp, 1))
    starts = np.sort(rng.integers(lo, max(span_bp, 1), size=n_cand))
    lens = rng.lognormal(np.log(read_len_mean), read_len_sigma, size=n_cand)
    cov = np.zeros(C, dtype=np.int64)
    reads = []  # (first_col, last_col, side)
    for s, L in zip(starts, lens):
        first = int(np.searchsorted(positions, max(s, 0), side="left"))
        last = int(np.searchsorted(positions, s + L, side="right")) - 1
        if last - first + 1 < 2 or first >= C:
            continue
        if cov[first:last + 1].max(initial=0) >= max_coverage:
            continue
        cov[first:last + 1] += 1
        reads.append((first, last, int(rng.integers(2))))
    reads.sort(key=lambda t: t[0])  # ReadSet order: by first position (ties: as generated)
    N = len(reads)
    read_tag = np.full(N, -1, dtype=np.int8)
    if tags == "all":
        read_tag[:] = [r[2] for r in reads]
    elif tags == "mixed":
        sel = rng.random(N) < 0.5
        read_tag[sel] = np.array([r[2] for r in reads], dtype=np.int8)[sel]
    elif tags != "none":
        raise ValueError(tags)
    offs = [0]
    e_col, e_all, e_soff, e_sc = [], [], [0], []
    for first, last, side in reads:
        cols = np.arange(first, last + 1)
        if len(cols) > 2 and gap_frac > 0:
            keep = rng.random(len(cols)) >= gap_frac
            keep[0] = keep[-1] = True
            cols = cols[keep]
        for c in cols:
            n = int(n_alleles[c])
            t = int(truth[side, c])
            d = rng.poisson(3, size=n) + 1
            if rng.random() < accuracy:
                d[t] = 0
            else:
                d[int(rng.integers(n))] = 0
            sc = -d.astype(np.float64)
            obs = int(np.argmax(sc))
            if unknown_allele_frac > 0 and rng.random() < unknown_allele_frac:
                obs = -1
            e_col.append(int(c))
            e_all.append(obs)
            e_sc.extend(sc.tolist())
            e_soff.append(len(e_sc))
        offs.append(len(e_col))
    return Problem(
        positions=positions,
        n_alleles=n_alleles,
        allele_ref=panel,
        recomb=uniform_recombination_cost(positions),
        read_tag=read_tag,
        read_entry_offset=np.array(offs, dtype=np.uint32),
        entry_column=np.array(e_col, dtype=np.uint32),
        entry_allele=np.array(e_all, dtype=np.int32),
        entry_score_offset=np.array(e_soff, dtype=np.uint64),
        entry_scores=np.array(e_sc, dtype=np.float64),
        reg_const=reg_const,
        base_const=base_const,
        meta=dict(seed=seed, tags=tags, coverage=coverage, max_coverage=max_coverage),
    )


# ---- the BASELINE.json configs, at the sizes this repo actually runs ----------
def config1(seed=1, tags="none", n_columns=5000):
    """configs[0]: 1 Mb, 10-haplotype panel, ~5k bubbles, 10x HiFi (CPU-runnable)."""
    return make_problem(n_columns, 10, region_bp=1_000_000 * n_columns // 5000, coverage=10.0,
                        max_coverage=15, tags=tags, seed=seed)


def config2_window(n_columns=64, seed=2, tags="none", max_coverage=15, n_haps=88):
    """configs[1] shape: 88-haplotype panel, 30x HiFi capped at max-coverage 15,
    variant density of 516k sites / 64 Mb; a window of ``n_columns`` columns."""
    return make_problem(n_columns, n_haps, region_bp=124 * n_columns, coverage=30.0,
                        max_coverage=max_coverage, tags=tags, window=True, seed=seed)


def config4_window(n_columns=16, seed=4, n_haps=10, max_coverage=20):
    """configs[3] shape: 60x ONT, max-coverage 20 (2^20 bipartitions per column)."""
    return make_problem(n_columns, n_haps, region_bp=124 * n_columns, coverage=60.0,
                        max_coverage=max_coverage, read_len_mean=30000.0, read_len_sigma=0.6,
                        accuracy=0.93, tags="none", window=True, seed=seed)


def config5_window(n_columns=16, seed=5, n_haps=400, max_coverage=6):
    """configs[4] shape: 400-haplotype panel, multiallelic bubbles (3..16 alleles everywhere)."""
    return make_problem(n_columns, n_haps, region_bp=124 * n_columns, coverage=15.0,
                        max_coverage=max_coverage, multiallelic_frac=1.0, max_alleles=16,
                        tags="none", window=True, seed=seed)
