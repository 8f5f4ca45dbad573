"""Full-length bookkeeping for the benchmark: what the checkpoint schedule of a whole chromosome costs, and a
state budget under which a short window pays the same recompute per column.

configs[1] is 516k columns of up to 1 GB each against 180 GB of HBM: the backward columns are produced about 2.8
times each (giggles_b200/csrc/schedule.cpp), not once.  A benchmark window of a few hundred columns under the default
budget would be swept with a far milder schedule; :func:`budget_for_ratio` finds the budget that makes the window's
schedule recompute as much per column as the full chromosome's does, so that one step of the window costs what the
same columns cost inside the full-length run.  Host only (``gg_schedule_dry_run``).
"""
from __future__ import annotations

import numpy as np

from . import capi


def coverage_profile(n_columns: int, seed: int = 2, max_coverage: int = 15, coverage: float = 30.0,
                     read_len_mean: float = 15000.0, read_len_sigma: float = 0.2, bp_per_column: float = 124.0,
                     window: bool = True) -> np.ndarray:
    """u_c (untagged active reads per column) of a synthetic chain generated like ``synth.make_problem`` does —
    same positions, read starts / lengths and greedy coverage cap — without panel, scores or entries, so that
    chromosome-length chains (516k columns) take seconds."""
    rng = np.random.default_rng(seed)
    C = int(n_columns)
    gaps = np.maximum(1, np.round(rng.exponential(bp_per_column, size=C))).astype(np.int64)
    positions = np.cumsum(gaps)
    span_bp = int(positions[-1])
    n_cand = int(np.ceil(coverage * span_bp / read_len_mean * 1.3))
    lo = -int(2 * read_len_mean) if window else 0
    n_cand = int(n_cand * (span_bp - lo) / span_bp)
    starts = np.sort(rng.integers(lo, span_bp, size=n_cand))
    lens = rng.lognormal(np.log(read_len_mean), read_len_sigma, size=n_cand)
    firsts = np.searchsorted(positions, np.maximum(starts, 0), side="left")
    lasts = np.searchsorted(positions, starts + lens, side="right") - 1
    cov = np.zeros(C, dtype=np.int64)
    for first, last in zip(firsts.tolist(), lasts.tolist()):
        if last - first + 1 < 2 or first >= C:
            continue
        if cov[first:last + 1].max(initial=0) >= max_coverage:
            continue
        cov[first:last + 1] += 1
    return cov


def column_bytes(u: np.ndarray, n_haps: int, n_alleles=2, elem: int = 4):
    """(column bytes, F/G table bytes) per column in the device layout (block record R^2 + 2R + 4 elements)."""
    R = int(n_haps)
    Rl = R + (R & 1)                       # odd panels carry a phantom haplotype
    if Rl > 256:
        Rl = (Rl + 3) // 4 * 4
    bs = (Rl * Rl + 2 * Rl + 4 + 3) // 4 * 4
    u = np.asarray(u, dtype=np.int64)
    nA = np.broadcast_to(np.asarray(n_alleles, dtype=np.int64), u.shape)
    cb = (bs * elem * (1 << u)).astype(np.uint64)
    fb = (2 * (nA + 1) * elem * (1 << u)).astype(np.uint64)
    return cb, fb


def schedule_cost(cb, fb, budget: int) -> dict:
    """Dry run + the two ratios that matter: backward launches per column and backward column bytes per byte of
    column (1.0 = every backward column produced once)."""
    r = capi.schedule_dry_run(cb, fb, int(budget))
    n = max(len(cb), 1)
    r["bwd_per_column"] = r["n_bwd"] / n
    r["bwd_bytes_ratio"] = r["bwd_col_bytes"] / max(float(np.sum(cb[:-1], dtype=np.float64)), 1.0) if n > 1 else 1.0
    return r


def budget_for_ratio(cb, fb, target_ratio: float, max_budget: int) -> tuple[int, dict]:
    """The state budget (bytes) under which the schedule of these columns produces at least ``target_ratio`` bytes of
    backward column per byte of column, and as little above it as possible; scanned on a grid of quarter columns.
    Returns (budget, schedule_cost).  If even the tightest feasible budget stays below the target, that one."""
    cb = np.asarray(cb, dtype=np.uint64)
    fb = np.asarray(fb, dtype=np.uint64)
    step = max(int(cb.max()) // 4, 1 << 20)
    best = None          # (ratio, budget, cost) with ratio >= target, smallest ratio
    below = None         # closest from below
    b = 3 * int(cb.max())
    top = min(int(max_budget), int(cb.sum() + fb.sum()) + 6 * int(cb.max()))
    while b <= top:
        try:
            c = schedule_cost(cb, fb, b)
        except capi.GGError:
            b += step
            continue
        r = c["bwd_bytes_ratio"]
        if r >= target_ratio:
            if best is None or r < best[0]:
                best = (r, b, c)
        elif below is None or r > below[0]:
            below = (r, b, c)
        if r <= 1.0 + 1e-9:
            break
        b += step
    pick = best or below
    if pick is None:
        raise capi.GGError(capi.GG_ERR_NOMEM, "no feasible budget found")
    return pick[1], pick[2]
