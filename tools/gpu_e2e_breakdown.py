import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from giggles_b200 import capi
from giggles_b200.synth import config2_window
p = config2_window(256, seed=2, max_coverage=15)
t = time.perf_counter(); cp = capi.CProblem(p); print("CProblem (entry emissions)", time.perf_counter() - t)
for it in range(4):
    t0 = time.perf_counter(); h = capi.Hmm(cp, device=0); t1 = time.perf_counter()
    h.sweep(); t2 = time.perf_counter()
    r = h.fetch(); t3 = time.perf_counter()
    h.close(); t4 = time.perf_counter()
    print(f"iter {it}: create {1e3*(t1-t0):.1f} ms  sweep {1e3*(t2-t1):.1f} ms  fetch {1e3*(t3-t2):.1f} ms  destroy {1e3*(t4-t3):.1f} ms")
t0 = time.perf_counter(); r = capi.hmm_run(cp, device=0); print("hmm_run", 1e3*(time.perf_counter()-t0), "ms")
t0 = time.perf_counter(); r = capi.hmm_run(cp, device=0); print("hmm_run", 1e3*(time.perf_counter()-t0), "ms")
