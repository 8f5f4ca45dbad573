"""One-off: the pipelined step kernel on FULL-SIZE columns of the bench shape (R = 88, 2^15 bipartition blocks, 1.04 GB
per column) against the CPU oracle — five columns, two reads starting and two ending inside the window (the oracle's cost doubles
with every read that starts or ends between two columns, hence only a few).  The oracle needs about 25 s per such
column on one core; the result goes to profiles/ (this is the shape bench.py times)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import oracle  # noqa: E402
from giggles_b200 import capi  # noqa: E402
from giggles_b200.synth import make_problem  # noqa: E402

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1511
p = make_problem(5, 88, region_bp=5 * 124, coverage=40, max_coverage=15, read_len_mean=5000, read_len_sigma=0.6,
                 window=True, seed=seed, base_const=1.1, accuracy=0.8)    # weak evidence: unsaturated posteriors
u = p.untagged_per_column().tolist()
pv = capi.PlanView(p)
started = ended = 0
for c in range(1, p.n_columns):
    a, b = set(pv.untagged(c - 1)), set(pv.untagged(c))
    started += len(b - a)
    ended += len(a - b)
t = time.time(); got = {kv: capi.hmm_run(p, device=0, kernel_variant=kv).values for kv in (0, 2, 3)}; t_gpu = time.time() - t
t = time.time(); want = oracle.oracle_run(p); t_cpu = time.time() - t
out = {"n_haps": 88, "untagged_per_column": u, "reads_started_inside": started, "reads_ended_inside": ended,
       "state_columns": p.state_columns(), "oracle_seconds_1_core": round(t_cpu, 1), "gpu_seconds_3_variants_e2e": round(t_gpu, 2)}
big = want >= 1e-3
for kv, g in got.items():
    out[f"variant{kv}_max_rel_err_p_ge_1e-3"] = float(np.max(np.abs(g[big] - want[big]) / want[big]))
    out[f"variant{kv}_max_abs_err_p_lt_1e-3"] = float(np.max(np.abs(g[~big] - want[~big]))) if (~big).any() else 0.0
out["posteriors_strictly_inside_1e-4_0.999"] = float(np.mean((want > 1e-4) & (want < 0.999)))
out["pass"] = all(out[f"variant{kv}_max_rel_err_p_ge_1e-3"] <= 1e-5 and out[f"variant{kv}_max_abs_err_p_lt_1e-3"] <= 2e-6 for kv in got)
print(json.dumps(out))
