"""configs[0] at FULL size (5,000 variant sites, 10-haplotype panel, 10x untagged): CUDA path vs the CPU oracle.
Prints the parity summary that DESIGN.md quotes.  ~3 minutes of CPU for the oracle."""
import sys, os, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle
from giggles_b200 import capi
from giggles_b200.synth import config1
from conftest import py_determine_genotype, py_gq

p = config1(seed=1, n_columns=int(sys.argv[1]) if len(sys.argv) > 1 else 5000)
t = time.time(); r = capi.hmm_run(p, device=0); t_gpu = time.time() - t
t = time.time(); want = oracle.oracle_run(p); t_cpu = time.time() - t
got = r.values
off = p.gl_offsets()
calls = r.call(0.0)
big = want >= 1e-3
rel = float(np.max(np.abs(got[big] - want[big]) / want[big]))
ab = float(np.max(np.abs(got[~big] - want[~big])))
same = ties = gq_off = 0
for c in range(p.n_columns):
    l = want[off[c]:off[c + 1]].tolist()
    srt = sorted(l)
    idx = py_determine_genotype(l)
    if abs(srt[-1] - srt[-2]) < 1e-5:
        ties += 1
        continue
    same += int(calls.gt_index[c] == idx)
    if idx >= 0 and py_gq(l, idx) < 60 and abs(int(calls.gq[c]) - py_gq(l, idx)) > 1:
        gq_off += 1
print(json.dumps({"columns": p.n_columns, "max_untagged": int(p.untagged_per_column().max()), "state_columns": p.state_columns(),
                  "max_rel_err_p_ge_1e-3": rel, "max_abs_err_p_lt_1e-3": ab, "same_GT": same, "near_ties_skipped": ties,
                  "GQ_off_by_more_than_1": gq_off, "gpu_seconds_e2e": round(t_gpu, 3), "oracle_seconds_1_core": round(t_cpu, 1)}))
