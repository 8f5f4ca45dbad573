"""Small problems that reach every kernel family (emission, generic step, warp-per-block, pipelined bulk-copy step
incl. two-block reduce and row tiles, chain) in both precisions and all kernel variants, with GG_POISON_ARENA=1 so
that a read of memory the sweep did not write shows up as NaN.  (compute-sanitizer is closed on the GPU pool; this is
the bounded stand-in.)  Results are checked for finiteness only; parity is the job of tests/."""
import os
os.environ.setdefault("GG_POISON_ARENA", "1")
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from giggles_b200 import capi
from giggles_b200.synth import make_problem

cases = [
    dict(n_columns=12, n_haps=4, coverage=3, max_coverage=3, read_len_mean=1500),
    dict(n_columns=20, n_haps=10, coverage=6, max_coverage=6, read_len_mean=2500, multiallelic_frac=0.4, max_alleles=6),
    dict(n_columns=40, n_haps=8, coverage=6, max_coverage=6, read_len_mean=2000, tags="all"),
    dict(n_columns=16, n_haps=88, coverage=8, max_coverage=6, read_len_mean=3000, tags="mixed", window=True, multiallelic_frac=0.3),
    dict(n_columns=12, n_haps=88, coverage=6, max_coverage=5, read_len_mean=800, missing_frac=0.05),
    dict(n_columns=10, n_haps=33, coverage=5, max_coverage=4, read_len_mean=1500),
    dict(n_columns=6, n_haps=200, coverage=3, max_coverage=2, read_len_mean=1500, multiallelic_frac=1.0, max_alleles=16, window=True),
    dict(n_columns=6, n_haps=400, coverage=4, max_coverage=2, read_len_mean=700, multiallelic_frac=1.0, max_alleles=16, window=True),
    dict(n_columns=8, n_haps=94, coverage=8, max_coverage=5, read_len_mean=600, tags="mixed", window=True),
]
variants = [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else "0,1,2,3,4,5".split(","))]
n = 0
for i, kw in enumerate(cases):
    p = make_problem(seed=300 + i, **kw)
    for kv in variants:
        for prec in (0, 1):
            v = capi.hmm_run(p, device=0, precision=prec, kernel_variant=kv).values
            assert np.isfinite(v).all(), (i, kv, prec)
            n += 1
print("sanitize_cases: ran", n, "sweeps over", len(cases), "problems")
