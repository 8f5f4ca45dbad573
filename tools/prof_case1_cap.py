"""Development helper (GPU box): short chain of R = 10 with every column near the coverage cap (argv[1], default 15), for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from giggles_b200 import capi
from giggles_b200.synth import make_problem
cap = int(sys.argv[1]) if len(sys.argv) > 1 else 15
n = int(sys.argv[2]) if len(sys.argv) > 2 else 300
p = make_problem(n, 10, region_bp=200 * n, coverage=2.0 * cap, max_coverage=cap, seed=5)
h = capi.Hmm(p, device=0)
h.sweep(); h.sweep(True)
s = h.stats()
print({k: s[k] for k in ("n_columns", "state_columns", "max_untagged", "n_launches", "last_sweep_ms", "last_fwd_ms", "last_bwd_ms", "last_emit_ms")})
h.close()
