"""Development helper (GPU box): independent all-haplotagged chains run concurrently from host threads (each gg_hmm_run
has its own stream; a u=0 chain is one persistent CTA, so one chain uses one of 148 SMs)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from concurrent.futures import ThreadPoolExecutor
from giggles_b200 import capi, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
cols = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
probs = [capi.CProblem(synth.config2_window(cols, seed=10 + i, tags="all")) for i in range(n)]
capi.hmm_run(probs[0], device=0)
t = time.perf_counter(); r1 = [capi.hmm_run(p, device=0) for p in probs]; t1 = time.perf_counter() - t
print(f"{n} chains x {cols} tagged columns: sequential {1e3*t1:.1f} ms = {n*cols/t1:,.0f} variants/s", flush=True)
capi.hmm_run_batch(probs, device=0)
t = time.perf_counter(); rb = capi.hmm_run_batch(probs, device=0); tb = time.perf_counter() - t
print(f"  gg_hmm_run_batch {1e3*tb:.1f} ms = {n*cols/tb:,.0f} variants/s, identical {all(np.array_equal(a.values, b.values) for a, b in zip(r1, rb))}", flush=True)
for w in (2, 4, 8, 16, 24):
    capi.hmm_run_many(probs, device=0, workers=w)      # warm the arena pool for this width
    t = time.perf_counter(); r2 = capi.hmm_run_many(probs, device=0, workers=w); t2 = time.perf_counter() - t
    same = all(np.array_equal(a.values, b.values) for a, b in zip(r1, r2))
    print(f"  {w} host threads: {1e3*t2:.1f} ms = {n*cols/t2:,.0f} variants/s, identical {same}", flush=True)
