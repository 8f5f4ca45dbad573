import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from giggles_b200 import capi
from giggles_b200 import synth
works = {"config2": lambda: synth.config2_window(256, seed=2, max_coverage=15),
         "tagged": lambda: synth.config2_window(20000, seed=2, tags="all"),
         "config5": lambda: synth.config5_window(48, seed=2, n_haps=400, max_coverage=8),
         "config1": lambda: synth.config1(seed=2, n_columns=5000)}
for name in sys.argv[1:] or list(works):
    p = works[name]()
    t = time.perf_counter(); cp = capi.CProblem(p); tc = time.perf_counter() - t
    for it in range(int(os.environ.get('GG_ITERS', '3'))):
        t0 = time.perf_counter(); h = capi.Hmm(cp, device=0); t1 = time.perf_counter()
        st = h.stats()
        h.sweep(); t2 = time.perf_counter()
        r = h.fetch(); t3 = time.perf_counter()
        h.close(); t4 = time.perf_counter()
        print(f"{name} iter {it}: reserved {st['state_bytes_reserved']/1e9:.2f} GB bwd {st['n_backward_steps']} CProblem {1e3*tc:.1f} create {1e3*(t1-t0):.1f} ms (plan {st['create_plan_ms']:.1f}, total {st['create_total_ms']:.1f})  sweep {1e3*(t2-t1):.1f} ms  fetch {1e3*(t3-t2):.1f} ms  destroy {1e3*(t4-t3):.1f} ms", flush=True)
