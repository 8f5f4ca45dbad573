"""Development helper (GPU box): one short sweep of a configs[4]-shaped window (400 haplotypes, multi-allelic), for ncu."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from giggles_b200 import capi
from giggles_b200.synth import config5_window
mc = int(sys.argv[1]) if len(sys.argv) > 1 else 12
ncol = int(sys.argv[2]) if len(sys.argv) > 2 else 8
kv = int(sys.argv[3]) if len(sys.argv) > 3 else 0
p = config5_window(ncol, seed=2, n_haps=400, max_coverage=mc)
h = capi.Hmm(p, device=0, kernel_variant=kv)
h.sweep()
h.sweep(True)
s = h.stats()
print({k: s[k] for k in ("n_columns", "state_columns", "max_untagged", "n_launches", "last_sweep_ms", "last_fwd_ms", "last_bwd_ms")})
h.close()
