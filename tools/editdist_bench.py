"""Throughput of the batched device edit distance (gg_edit_distance_batch, H2D / D2H inside) against the single-core C
restatement of the reference's edit_distance (oracle/edit_oracle.c), on the shapes the allele detection produces:
a read window against every allele of a variant (giggles/variants.py:229-239).  Results are compared pair by pair."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import oracle
from giggles_b200 import align

rng = np.random.default_rng(11)


def batch(n, lo, hi, err=0.1):
    pairs = []
    for _ in range(n):
        L = int(rng.integers(lo, hi + 1))
        s = rng.choice(list(b"ACGT"), size=L).astype(np.uint8)
        t = s.copy()
        k = rng.random(L) < err                      # substitutions, then a few indels
        t[k] = rng.choice(list(b"ACGT"), size=int(k.sum()))
        t = np.delete(t, rng.integers(0, L, size=max(1, L // 50)))
        pairs.append((s.tobytes(), t.tobytes()))
    return pairs


for name, n, lo, hi, md in (("allele windows 30..64 bp", 400000, 30, 64, -1), ("windows 100..300 bp", 200000, 100, 300, -1),
                            ("SV alleles 1..3 kbp", 20000, 1000, 3000, -1), ("windows 100..300 bp, banded maxdiff 20", 200000, 100, 300, 20)):
    pairs = batch(n, lo, hi)
    align.edit_distance_batch(pairs[:1000], md)      # warm
    t0 = time.time(); packed = align.pack_pairs(pairs); t_pack = time.time() - t0
    t0 = time.time(); got = align.edit_distance_packed(packed, md); t_gpu = time.time() - t0
    m = min(n, 20000)
    t0 = time.time(); want = np.array([oracle.oracle_edit_distance(s, t, md) for s, t in pairs[:m]]); t_cpu = (time.time() - t0) * n / m
    assert np.array_equal(got[:m], want)
    cells = float(sum(len(s) * len(t) for s, t in pairs))
    print(f"{name}: {n} pairs, {cells:.3g} DP cells: gg_edit_distance_batch {n / t_gpu:,.0f} pairs/s ({cells / t_gpu / 1e9:.1f} Gcell/s; H2D / D2H inside, Python packing {t_pack:.2f} s outside), "
          f"1-core C restatement {n / t_cpu:,.0f} pairs/s (extrapolated from {m}) -> {t_cpu / t_gpu:.0f}x; identical on the {m} compared")
