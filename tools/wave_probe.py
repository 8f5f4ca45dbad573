"""Per-step time of the wave kernel against the column size: chains of R = 10 with the coverage cap (= max u) varied,
so that the fixed cost per step (grid barrier, argument staging, scale reduction) separates from the per-block cost."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from giggles_b200 import capi
from giggles_b200.synth import make_problem

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
for cap in (1, 4, 8, 11, 13, 15):
    p = make_problem(n, 10, region_bp=200 * n, coverage=2.0 * cap, max_coverage=cap, seed=5)
    u = p.untagged_per_column()
    h = capi.Hmm(p, device=0)
    h.sweep(); h.sweep(True)
    s = h.stats()
    print(f"cap {cap:2d} mean 2^u {float(np.mean(2.0 ** u)):9.1f} blocks  fwd {1e3 * s['last_fwd_ms'] / n:6.2f} us/step  "
          f"bwd {1e3 * s['last_bwd_ms'] / n:6.2f} us/step  launches {s['n_launches']}")
    h.close()
