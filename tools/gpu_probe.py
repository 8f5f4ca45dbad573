"""Development helper (GPU box): time sweeps of config-2-shaped windows, print GB/s per kernel class."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from giggles_b200 import capi
from giggles_b200.synth import config2_window, config1

def run(name, p, budget=0, reps=3, kv=0):
    t = time.time(); h = capi.Hmm(p, device=0, budget=budget, kernel_variant=kv); tc = time.time() - t
    st = h.stats()
    for _ in range(2): h.sweep()
    best = 1e9
    for _ in range(reps):
        h.sweep(False); best = min(best, h.stats()['last_sweep_ms'])
    h.sweep(True)
    s = h.stats()
    S = st['state_columns']
    print(f"{name}: C={st['n_columns']} maxu={st['max_untagged']} states={S:.3e} arena={st['state_bytes_reserved']/2**30:.2f} GiB levels={st['checkpoint_levels']} "
          f"launches={s['n_launches']} sweep={best:.3f} ms  alg GB/s={20*S/best/1e6:.1f} sched GB/s={st['scheduled_bytes']/best/1e6:.1f} "
          f"fwd={s['last_fwd_ms']:.3f} bwd={s['last_bwd_ms']:.3f} emit={s['last_emit_ms']:.3f} ms create={tc*1e3:.0f} ms cols/s={st['n_columns']/best*1e3:.0f}", flush=True)
    h.close()

which = sys.argv[1:] or ["small", "mid", "big"]
if "small" in which:
    run("cfg1 R=10 C=2000", config1(n_columns=2000))
    run("cfg2 tagged R=88 C=2000", config2_window(2000, tags="all"))
if "tagged" in which:
    for _ in range(2):
        run("tagged R=88 C=4000 fast", config2_window(4000, tags="all"))
        run("tagged R=88 C=4000 smem-posterior", config2_window(4000, tags="all"), kv=6)
if "mid" in which:
    run("cfg2 mc=10 C=64", config2_window(64, max_coverage=10))
    run("cfg2 mc=12 C=64", config2_window(64, max_coverage=12))
if "r400" in which:
    from giggles_b200.synth import config5_window
    run("cfg5 R=400 mc=6 C=24", config5_window(24, max_coverage=6))
    run("cfg5 R=400 mc=6 C=24 generic", config5_window(24, max_coverage=6), kv=1)
    run("R=256 biallelic mc=8 C=24", config2_window(24, max_coverage=8, n_haps=256))
if "r94" in which:
    run("R=94 mc=15 C=48", config2_window(48, max_coverage=15, n_haps=94))
    run("R=94 mc=15 C=48 generic", config2_window(48, max_coverage=15, n_haps=94), kv=1)
if "cw16" in which:
    run("cfg2 mc=15 C=160 auto", config2_window(160, max_coverage=15), reps=2)
    run("cfg2 mc=15 C=160 one CTA x 16 warps", config2_window(160, max_coverage=15), reps=2, kv=2)
if "big" in which:
    run("cfg2 mc=15 C=48", config2_window(48, max_coverage=15))
    run("cfg2 mc=15 C=160 (checkpointed)", config2_window(160, max_coverage=15), reps=1)
