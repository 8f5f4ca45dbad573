"""Development helper (GPU box): CUDA path vs CPU oracle over a spread of seeded problems."""
import sys, os, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import oracle
from giggles_b200 import capi
from giggles_b200.synth import make_problem

cases = [
    dict(n_columns=1, n_haps=4, coverage=3, max_coverage=3, read_len_mean=800),
    dict(n_columns=12, n_haps=4, coverage=3, max_coverage=3, read_len_mean=1500),
    dict(n_columns=40, n_haps=6, coverage=5, max_coverage=6, read_len_mean=2000, tags="none"),
    dict(n_columns=40, n_haps=5, coverage=5, max_coverage=6, read_len_mean=2000, tags="mixed"),
    dict(n_columns=40, n_haps=8, coverage=6, max_coverage=6, read_len_mean=2000, tags="all"),
    dict(n_columns=60, n_haps=10, coverage=8, max_coverage=8, read_len_mean=2500, multiallelic_frac=0.4, max_alleles=6),
    dict(n_columns=30, n_haps=7, coverage=6, max_coverage=7, read_len_mean=2500, unknown_allele_frac=0.1, missing_frac=0.05),
    dict(n_columns=24, n_haps=88, coverage=4, max_coverage=4, read_len_mean=2000, window=True),
    dict(n_columns=16, n_haps=88, coverage=8, max_coverage=6, read_len_mean=3000, tags="mixed", window=True, multiallelic_frac=0.3),
    dict(n_columns=10, n_haps=36, coverage=6, max_coverage=5, read_len_mean=1500, window=True),
    dict(n_columns=8, n_haps=200, coverage=3, max_coverage=2, read_len_mean=1500, multiallelic_frac=1.0, max_alleles=16, window=True),
    dict(n_columns=20, n_haps=33, coverage=5, max_coverage=4, read_len_mean=1500),
    dict(n_columns=40, n_haps=12, coverage=10, max_coverage=9, read_len_mean=1200, gap_frac=0.2),
    dict(n_columns=30, n_haps=64, coverage=8, max_coverage=6, read_len_mean=900, multiallelic_frac=0.5, max_alleles=16),
    dict(n_columns=30, n_haps=128, coverage=6, max_coverage=4, read_len_mean=900, tags="mixed"),
    dict(n_columns=30, n_haps=88, coverage=10, max_coverage=7, read_len_mean=600, missing_frac=0.05),
    dict(n_columns=10, n_haps=400, coverage=6, max_coverage=3, read_len_mean=700, multiallelic_frac=1.0, max_alleles=16, window=True),
    dict(n_columns=12, n_haps=256, coverage=8, max_coverage=4, read_len_mean=500, tags="mixed", multiallelic_frac=0.3),
    dict(n_columns=12, n_haps=132, coverage=8, max_coverage=4, read_len_mean=500, missing_frac=0.03),
    dict(n_columns=8, n_haps=512, coverage=4, max_coverage=2, read_len_mean=600),
    dict(n_columns=30, n_haps=94, coverage=10, max_coverage=6, read_len_mean=800, multiallelic_frac=0.3, max_alleles=8),
    dict(n_columns=30, n_haps=94, coverage=8, max_coverage=5, read_len_mean=600, tags="mixed", window=True),
    dict(n_columns=20, n_haps=34, coverage=8, max_coverage=6, read_len_mean=800),
    dict(n_columns=20, n_haps=66, coverage=8, max_coverage=6, read_len_mean=800, missing_frac=0.05),
    dict(n_columns=12, n_haps=130, coverage=8, max_coverage=4, read_len_mean=500, multiallelic_frac=0.5),
    dict(n_columns=10, n_haps=254, coverage=6, max_coverage=3, read_len_mean=500, tags="mixed"),
]
worst = 0.0
for i, kw in enumerate(cases):
    p = make_problem(seed=100 + i, **kw)
    want = oracle.oracle_run(p)
    for prec, tol, kv in ((0, 2e-5, 0), (0, 2e-5, 1), (1, 1e-11, 0)):
        t = time.time()
        try:
            got = capi.hmm_run(p, device=0, precision=prec, kernel_variant=kv).values
        except Exception as e:
            print(f"case {i} prec {prec}: ERROR {e}")
            continue
        dt = time.time() - t
        err = np.abs(got - want)
        bad = np.isnan(got).sum()
        tag = "OK " if (err.max() < tol and not bad) else "BAD"
        worst = max(worst, err.max()) if prec == 0 else worst
        print(f"{tag} case {i} R={p.n_haps} C={p.n_columns} maxu={p.untagged_per_column().max()} prec={prec} kv={kv} "
              f"maxerr={err.max():.3e} nan={bad} at={int(err.argmax())} {dt*1e3:.1f} ms")
        if tag == "BAD":
            off = p.gl_offsets()
            nanc = [c for c in range(p.n_columns) if np.isnan(got[off[c]:off[c+1]]).any()]
            print("   NaN columns:", nanc[:50], "of", p.n_columns)
            c = int(np.searchsorted(off, err.argmax(), side='right') - 1)
            print("   first bad column", c, "got", got[off[c]:off[c+1]], "want", want[off[c]:off[c+1]])
# budget variations must give bit-identical results
p = make_problem(64, 12, coverage=6, max_coverage=6, read_len_mean=2500, tags="none", seed=5)
base = capi.hmm_run(p, device=0).values
h = capi.Hmm(p, device=0); st = h.stats(); h.close()
for frac in (0.6, 0.3, 0.15, 0.08):
    try:
        h = capi.Hmm(p, device=0, budget=int(st['state_bytes_reserved'] * frac))
        h.sweep(); v = h.fetch().values; s2 = h.stats(); h.close()
        print(f"budget x{frac}: levels={s2['checkpoint_levels']} bwd={s2['n_backward_steps']} identical={np.array_equal(v, base)} maxdiff={np.abs(v-base).max():.2e}")
    except Exception as e:
        print(f"budget x{frac}: {e}")
print("worst fp32 err", worst)
