"""Development helper (GPU box): achieved GB/s per launch category on the bench workload."""
import sys, os, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from giggles_b200 import capi
from giggles_b200.synth import config2_window
p = config2_window(256, seed=2, max_coverage=15)
pv = capi.PlanView(p)
u = p.untagged_per_column()
nsh = np.array([pv.shared_mask_prev(c)[1] for c in range(p.n_columns)])
nA = p.n_alleles
h = capi.Hmm(p, device=0)
for _ in range(2): h.sweep()
h.sweep(True)
k, c, ms = h.debug_op_times()
h.close()
R = p.n_haps
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for kind, col, t in zip(k, c, ms):
    if kind == 3:      # forward producing col
        nterm = (u[col - 1] - nsh[col]) if col > 0 else 0
        S = (1 << int(u[col])) * R * R
        byts = 4.0 * (2 * S + (S << int(nterm) if col > 0 else 0))
        cat = ("fwd", "nred=%d" % (1 << int(nterm)), "nA=%d" % (2 if nA[col] == 2 else 3))
    elif kind == 2:    # backward producing col from col+1
        nnew = u[col + 1] - nsh[col + 1]
        S = (1 << int(u[col])) * R * R
        Sin = (1 << int(u[col + 1])) * R * R
        byts = 4.0 * (S + Sin)
        cat = ("bwd", "nred=%d" % (1 << int(nnew)), "nA=%d" % (2 if nA[col] == 2 else 3))
    else:
        continue
    a = agg[cat]; a[0] += 1; a[1] += t; a[2] += byts
tot = sum(a[1] for a in agg.values())
for cat, (n, t, b) in sorted(agg.items()):
    print(f"{cat}: {n:4d} launches {t:8.2f} ms ({100*t/tot:4.1f}%)  actual bytes moved {b/1e9:8.1f} GB -> {b/t/1e6:7.0f} GB/s")
