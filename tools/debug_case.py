import sys, os
os.environ["GG_POISON_ARENA"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import oracle
from giggles_b200 import capi
from giggles_b200.synth import make_problem
kw = dict(n_columns=40, n_haps=12, coverage=10, max_coverage=9, read_len_mean=1200, gap_frac=0.2)
p = make_problem(seed=112, **kw)
want = oracle.oracle_run(p)
off = p.gl_offsets()
u = p.untagged_per_column()
pv = capi.PlanView(p)
nsh = [pv.shared_mask_prev(c)[1] for c in range(p.n_columns)]
for kv in (0, 2, 1):
    h = capi.Hmm(p, device=0, kernel_variant=kv)
    h.sweep()
    got = h.fetch().values
    a, b = h.debug_scales()
    h.close()
    bad = [c for c in range(p.n_columns) if np.isnan(got[off[c]:off[c+1]]).any() or np.abs(got[off[c]:off[c+1]] - want[off[c]:off[c+1]]).max() > 1e-5]
    print("kv", kv, "bad columns", bad)
    print("  alpha scale nan at", np.where(~np.isfinite(a))[0].tolist()[:10], " beta scale nan at", np.where(~np.isfinite(b))[0].tolist()[-10:])
print("u per column", u.tolist())
print("n_sh per column", nsh)
