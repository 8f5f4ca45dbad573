"""Development helper (GPU box): randomized parity campaign, CUDA path (all kernel families) vs the CPU oracle."""
import sys, os, time
os.environ.setdefault("GG_POISON_ARENA", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import oracle
from giggles_b200 import capi
from giggles_b200.synth import make_problem

budget_s = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
t0 = time.time()
n = bad = 0
worst = 0.0
while time.time() - t0 < budget_s:
    R = int(rng.choice([4, 5, 6, 7, 8, 9, 10, 12, 15, 16, 20, 24, 31, 32, 33, 34, 87, 95, 127, 129, 36, 48, 50, 64, 66, 88, 94, 96, 126, 128, 130, 132, 160, 200, 254, 256, 400]))
    cap_states = 3e6 if R <= 128 else 6e6
    max_cov = int(rng.integers(0, 9))
    while (1 << max_cov) * R * R > cap_states and max_cov > 0:
        max_cov -= 1
    C = int(rng.integers(2, 40 if R <= 128 else 12))
    tags = str(rng.choice(["none", "mixed", "all"]))
    kw = dict(n_columns=C, n_haps=R, coverage=float(rng.uniform(0.5, 14)), max_coverage=max(max_cov, 1),
              read_len_mean=float(rng.choice([400, 800, 1500, 3000])), tags=tags,
              multiallelic_frac=float(rng.choice([0.0, 0.1, 0.5, 1.0])), max_alleles=int(rng.choice([3, 4, 8, 16])),
              missing_frac=float(rng.choice([0.0, 0.005, 0.05])), gap_frac=float(rng.choice([0.0, 0.02, 0.3])),
              unknown_allele_frac=float(rng.choice([0.0, 0.1])), window=bool(rng.integers(2)), seed=int(rng.integers(1 << 30)))
    if max_cov == 0:
        kw["tags"] = "all"
    p = make_problem(**kw)
    want = oracle.oracle_run(p)
    if np.isnan(want).any():
        continue
    for kv in (0, 1, 2, 3, 4, 5):
        try:
            got = capi.hmm_run(p, device=0, kernel_variant=kv).values
        except Exception as e:
            print("ERROR", kv, kw, e, flush=True)
            bad += 1
            continue
        big = want >= 1e-3
        rel = np.abs(got[big] - want[big]) / want[big] if big.any() else np.zeros(1)
        ab = np.abs(got[~big] - want[~big]) if (~big).any() else np.zeros(1)
        ok = (not np.isnan(got).any()) and rel.max() <= 1e-5 and ab.max() <= 2e-6
        worst = max(worst, float(rel.max()))
        if not ok:
            bad += 1
            print("BAD", kv, kw, "rel", rel.max(), "abs", ab.max(), "nan", int(np.isnan(got).sum()), "maxu", int(p.untagged_per_column().max()), flush=True)
    n += 1
print(f"stress: {n} problems x 6 kernel variants, {bad} failures, worst rel err on p>=1e-3: {worst:.2e}, {time.time()-t0:.0f} s")
