"""Development helper (GPU box): short sweep of a config-1-shaped chain (R = 10), for ncu."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from giggles_b200 import capi
from giggles_b200.synth import config1
p = config1(n_columns=int(sys.argv[1]) if len(sys.argv) > 1 else 60)
h = capi.Hmm(p, device=0)
h.sweep(); h.sweep(True)
s = h.stats()
print({k: s[k] for k in ("n_columns", "state_columns", "max_untagged", "n_launches", "last_sweep_ms", "last_fwd_ms", "last_bwd_ms", "last_emit_ms")})
h.close()
