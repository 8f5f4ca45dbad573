#!/usr/bin/env python
"""Benchmark of the GenotypeHMM forward-backward (BASELINE.json metric: variants genotyped/s = HMM
columns/s; achieved HBM GB/s).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU GenotypeHMM on the host cores

Workload (default, `--workload genome`): BASELINE.json configs[1]'s column shape — 88-haplotype panel, 30x HiFi,
max-coverage 15, up to 2^15 x 88^2 states (1.04 GB) per column — as a fixed job of 24 independent chains whose
lengths follow the GRCh38 chromosomes (configs[2]'s partitioning), --total-columns variant sites in all.  One STEP is
one complete forward-backward of every chain: every backward column, every forward column, every posterior.

Recompute: at its real length (516k columns against 180 GB of HBM) the chromosome's checkpoint schedule produces
every backward column ~2.8 times (giggles_b200/csrc/schedule.cpp; tests/test_schedule.py asserts the bound).  A
short chain under the default budget would get away with ~1.1.  Every chain here therefore runs under the state budget
that makes ITS schedule recompute as many backward-column bytes per column as the full-length schedule does
(giggles_b200/fullscale.py); the `full_length` block of the JSON line states both.  Synthetic data, seeded.

N > 1: the chains are assigned to the ranks longest-first (giggles_b200/shard.py), each rank sweeps its share on its
GPU, the total is fixed ("strong" scaling); there is no data-path collective, the end-to-end arm adds the final
gather of posteriors to rank 0.  The side workloads (--workload config2/config1/tagged/config4/config5) run one
window per rank under the default budget ("weak").
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "variants_genotyped_per_s"
UNIT = "variants/s"
# GRCh38 chromosome lengths in Mb (chr1..22, X, Y): the chain lengths of the default workload
GRCH38_MB = [248, 242, 198, 190, 181, 171, 159, 145, 138, 133, 135, 133, 114, 107, 102, 90, 83, 80, 58, 64, 46, 50, 156, 57]
FULL_LENGTH_COLUMNS = 516000          # configs[1]: 516k variant sites over 64 Mb


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,"
         "clocks.mem,temperature.gpu")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power, mem, temp = [], [], set(), [], [], []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            try:
                mem.append(float(f[9])); temp.append(float(f[10]))
            except (ValueError, IndexError):
                pass
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "mem_mhz": float(np.median(mem)) if mem else None,
                "temp_c_max": max(temp) if temp else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# CPU baseline: the reference's own GenotypeHMM (oracle/_ref, compiled from the reference sources) or,
# where that library is absent, the C restatement.  Bounded sample of the same workload shape.
# ------------------------------------------------------------------------------------------------
def cpu_sample_problems(n, sample_cols, sample_cov, make):
    return [make(sample_cols, seed=1000 + i, max_coverage=sample_cov) for i in range(n)]


def run_cpu_sample(problems, cores):
    """Runs every problem once, `cores` at a time (the reference is single-threaded per chain; chains are
    independent, exactly like the GPU sharding).  Returns (state_columns/s, columns/s, seconds, kind)."""
    import concurrent.futures as cf
    import oracle
    kind = "reference" if oracle.have_reference() else "port"
    fn = oracle.reference_run if kind == "reference" else oracle.oracle_run
    if kind == "reference":
        oracle.ref_lib()
    else:
        oracle.oracle_lib()
    t0 = time.perf_counter()
    with cf.ThreadPoolExecutor(max_workers=cores) as ex:     # ctypes releases the GIL during the call
        list(ex.map(fn, problems))
    dt = time.perf_counter() - t0
    sc = sum(p.state_columns() for p in problems)
    cols = sum(p.n_columns for p in problems)
    return sc / dt, cols / dt, dt, kind


def cpu_sample_text(cores, args, probs, sc_s, spc, dt=None):
    return (f"EXTRAPOLATED, not same-work: the reference needs ~1e3 s per column of this workload, so it is timed on "
            f"{cores} chains x {args.cpu_sample_cols} columns, R={args.haps}, max-coverage {args.cpu_sample_cov} "
            f"({sum(p.state_columns() for p in probs):.3g} state-columns" + (f", {dt:.1f} s of CPU work" if dt else " per step") +
            f"); measured {sc_s:.4g} state-columns/s, divided by the workload's {spc:.4g} states per column")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="genome", choices=["genome", "config2", "config1", "tagged", "config4", "config5"],
                    help="genome (default, the headline): configs[1] columns as 24 chromosome-proportioned chains under the "
                         "full-length recompute; the others are side measurements of single windows (default budget)")
    ap.add_argument("--total-columns", type=int, default=1536, help="genome workload: variant sites over all 24 chains")
    ap.add_argument("--columns", type=int, default=None, help="side workloads: variant sites per window")
    ap.add_argument("--max-coverage", type=int, default=None, help="default: the workload's preset")
    ap.add_argument("--haps", type=int, default=None, help="default: the workload's preset")
    ap.add_argument("--recompute", default="full-length", choices=["full-length", "default"],
                    help="genome workload: full-length = per-chain budgets that reproduce the 516k-column schedule's "
                         "backward bytes per column; default = the library's default budget (94 %% of free HBM)")
    ap.add_argument("--cpu-sample-cols", type=int, default=240)   # ~10 s of CPU work per sample on 16 host cores
    ap.add_argument("--cpu-sample-cov", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    presets = {"genome": (None, 15, 88), "config2": (256, 15, 88), "config1": (5000, 15, 10), "tagged": (20000, 15, 88),
               "config4": (8, 20, 88), "config5": (24, 12, 400)}
    d_cols, d_cov, d_haps = presets[args.workload]
    if args.columns is None:
        args.columns = d_cols
    if args.max_coverage is None:
        args.max_coverage = d_cov
    if args.haps is None:
        args.haps = d_haps
    genome = args.workload == "genome"

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = os.cpu_count() or 1
    if args.impl == "reference" and rank != 0:
        return                                    # the CPU arm runs on rank 0 alone

    from giggles_b200 import fullscale, synth
    from giggles_b200.shard import chain_cost, lpt_assign

    def make_window(n_columns, seed=2, max_coverage=None):
        mc = args.max_coverage if max_coverage is None else max_coverage
        if args.workload == "config1":
            return synth.config1(seed=seed, n_columns=n_columns)
        if args.workload == "tagged":
            return synth.config2_window(n_columns, seed=seed, tags="all", n_haps=args.haps)
        if args.workload == "config5":
            return synth.config5_window(n_columns, seed=seed, n_haps=args.haps, max_coverage=mc)
        if args.workload == "config4":
            return synth.config4_window(n_columns, seed=seed, n_haps=args.haps, max_coverage=mc)
        return synth.config2_window(n_columns, seed=seed, max_coverage=mc, n_haps=args.haps)

    # ---------------------------------------------------------------- the job: a list of independent chains
    if genome:
        tot_mb = float(sum(GRCH38_MB))
        chain_cols = [max(8, int(round(args.total_columns * mb / tot_mb))) for mb in GRCH38_MB]
        chains = [make_window(n, seed=100 + i) for i, n in enumerate(chain_cols)]
        assign = lpt_assign([chain_cost(p) for p in chains], world)
        scaling = "strong"
    else:
        chains = [make_window(args.columns, seed=2) for _ in range(world)]     # same shape AND size on every rank
        assign = [[r] for r in range(world)]
        scaling = "weak"
    total_cols = float(sum(p.n_columns for p in chains))
    total_states = float(sum(p.state_columns() for p in chains))
    loads = [sum(chain_cost(chains[i]) for i in idx) for idx in assign]
    lpt = {"chains_per_rank": [len(idx) for idx in assign],
           "load_max_over_mean": (max(loads) / (sum(loads) / world)) if sum(loads) else 1.0,
           "largest_chain_share": max(chain_cost(p) for p in chains) / max(sum(loads), 1.0)}

    names = {"genome": f"configs[1] columns ({args.haps}-haplotype panel, 30x HiFi untagged, max-coverage {args.max_coverage}: <= 2^{args.max_coverage} x "
                       f"{args.haps}^2 states per column) as a fixed job of 24 chains with GRCh38 length ratios (configs[2]'s partitioning), "
                       f"{int(total_cols)} variant sites in all; " +
                       ("every chain under the state budget that reproduces the recompute of the full-length (516k-column) schedule"
                        if args.recompute == "full-length" else "default state budget (milder recompute than at full length)"),
             "config2": f"configs[1] chr20-scale window: {args.columns} of 516k variant sites, {args.haps}-haplotype panel, "
                        f"30x HiFi untagged, max-coverage {args.max_coverage}, default budget (NOT the full-length recompute)",
             "config1": f"configs[0] 1 Mb region: {args.columns} variant sites, 10-haplotype panel, 10x HiFi untagged",
             "tagged": f"configs[1] shape with every read haplotagged (the mode giggles runs in without --keep-untagged): "
                       f"{args.columns} variant sites, {args.haps}-haplotype panel, one {args.haps}^2 block per column",
             "config4": f"configs[3] shape: {args.columns} variant sites of 60x ONT, {args.haps}-haplotype panel, max-coverage "
                        f"{args.max_coverage} (2^{args.max_coverage} bipartition blocks = {(1 << args.max_coverage) * args.haps ** 2 * 4 / 2**30:.1f} GiB per column)",
             "config5": f"configs[4] shape: {args.columns} variant sites, {args.haps}-haplotype panel, 3..16 alleles per site, "
                        f"max-coverage {args.max_coverage}"}
    config = {"workload": names[args.workload],
              "columns_per_step": int(total_cols), "n_chains": len(chains), "n_haps": args.haps, "max_coverage": args.max_coverage,
              "parallelism": (f"{len(chains)} independent chains, longest-first onto {world} GPU(s), no data-path collective" if genome
                              else f"{world} independent chain(s), one per GPU"),
              "l2_policy": ("inputs larger than L2 (each column is up to 1 GB; L2 is 126 MB)" if args.workload in ("config2", "genome")
                            else "side measurement: columns may fit L2; not the headline number")}

    # ---------------------------------------------------------------- reference arm (CPU, rank 0 only)
    if args.impl == "reference":
        spc = total_states / total_cols
        probs = cpu_sample_problems(cores, args.cpu_sample_cols, args.cpu_sample_cov, make_window)
        vals = []
        for i in range(args.warmup + args.steps):
            sc_s, _, dt, kind = run_cpu_sample(probs, cores)
            if i >= args.warmup:
                vals.append((sc_s, dt))
        sc_s = sum(v[0] for v in vals) / len(vals)
        ms = 1e3 * sum(v[1] for v in vals) / len(vals)
        value = sc_s / spc
        sample = cpu_sample_text(cores, args, probs, sc_s, spc)
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "f80 (x87 long double)", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "state_columns_per_s": sc_s}))
        return

    # ---------------------------------------------------------------- our arm
    from giggles_b200 import capi
    from giggles_b200.shard import run_sharded
    if capi.lib().gg_device_count() < 1:
        raise SystemExit("bench.py: no CUDA device — the GenotypeHMM path has no CPU fallback")
    dist = None
    torch = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()

    # CPU baseline first (rank 0, N == 1 only), before the GPU is busy
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            spc = total_states / total_cols
            probs = cpu_sample_problems(cores, args.cpu_sample_cols, args.cpu_sample_cov, make_window)
            sc_s, _, dt, kind = run_cpu_sample(probs, cores)
            cpu_baseline = {"value": sc_s / spc, "unit": UNIT, "cores": cores, "kind": kind,
                            "sample": cpu_sample_text(cores, args, probs, sc_s, spc, dt), "state_columns_per_s": sc_s}
        except Exception as e:  # the baseline must never take the GPU number down with it
            cpu_baseline = {"value": None, "unit": UNIT, "cores": cores, "kind": "unavailable", "sample": str(e)}

    # what this device's HBM does today, measured the way the roofline denominator was (context for `roofline.frac`)
    box_copy_gbs = capi.lib().gg_device_copy_gbs(local_rank, 2 << 30, 5)

    # ---- the full-length schedule and the per-chain budgets that reproduce its recompute
    device_budget = int(capi.lib().gg_default_state_budget(local_rank))
    full_length = None
    budgets = {i: 0 for i in range(len(chains))}
    if genome:
        u_full = fullscale.coverage_profile(FULL_LENGTH_COLUMNS, max_coverage=args.max_coverage)
        cb, fb = fullscale.column_bytes(u_full, args.haps)
        fl = fullscale.schedule_cost(cb, fb, device_budget)
        full_length = {"columns": FULL_LENGTH_COLUMNS, "state_budget_bytes": device_budget, "checkpoint_levels": fl["levels"],
                       "backward_launches_per_column": fl["bwd_per_column"], "backward_bytes_per_column_byte": fl["bwd_bytes_ratio"],
                       "state_columns": float(np.sum(2.0 ** u_full) * args.haps ** 2)}
        if args.recompute == "full-length":
            for i in assign[rank]:
                p = chains[i]
                cbi, fbi = fullscale.column_bytes(p.untagged_per_column(), args.haps, p.n_alleles)
                budgets[i], _ = fullscale.budget_for_ratio(cbi, fbi, fl["bwd_bytes_ratio"], device_budget)

    mine = assign[rank]
    cps = {i: capi.CProblem(chains[i]) for i in mine}
    hs = {i: capi.Hmm(cps[i], device=local_rank, budget=budgets[i], shared_arena=genome) for i in mine}
    st0 = {i: hs[i].stats() for i in mine}
    for _ in range(args.warmup):
        for i in mine:
            hs[i].sweep()
    sampler = ClockSampler(local_rank)
    sampler.start()
    # ---- timed region: K steps over the resident chains, device time by the CUDA event pair gg_hmm_sweep puts around
    #      each sweep
    barrier()
    dev_ms, launches = 0.0, 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for i in mine:
            hs[i].sweep()              # synchronous: returns after the stream drained
            s = hs[i].stats()
            dev_ms += s["last_sweep_ms"]
            launches += s["n_launches"]
    my_wall = time.perf_counter() - t0
    barrier()
    wall = time.perf_counter() - t0
    # ---- the same K steps once more with a CUDA event pair around EVERY launch: the per-kernel times of `roofline`
    #      (the events cost a few per cent on the headline workload and much more where kernels last microseconds, so
    #      they are kept out of `value`)
    barrier()
    fwd_ms, bwd_ms, emit_ms, dev_ms_inst = 0.0, 0.0, 0.0, 0.0
    t0p = time.perf_counter()
    for _ in range(args.steps):
        for i in mine:
            hs[i].sweep(True)
            s = hs[i].stats()
            dev_ms_inst += s["last_sweep_ms"]
            fwd_ms += s["last_fwd_ms"]; bwd_ms += s["last_bwd_ms"]; emit_ms += s["last_emit_ms"]
    barrier()
    inst_wall = time.perf_counter() - t0p
    clocks = sampler.stop()
    res_dev = {i: hs[i].fetch().values for i in mine}
    for i in mine:
        hs[i].close()

    # ---- end to end through the public API with host buffers: per chain plan + H2D + sweep + D2H (gg_hmm_run), then the
    #      gather of every chain's posteriors to rank 0 (giggles_b200.shard.run_sharded)
    def run_one(p):
        i = next(k for k in mine if chains[k] is p)
        return capi.hmm_run(cps[i], device=local_rank, budget=budgets[i], shared_arena=genome).values

    dev = torch.device("cuda", local_rank) if dist is not None else None
    for _ in range(2):
        out = run_sharded(chains, run_one, rank, world, dist=dist, device=dev, assign=assign)   # warm (arena reuse, NCCL channels)
    barrier()
    e2e_steps = max(1, args.steps)
    e2e_step_ms = []
    t1 = time.perf_counter()
    for _ in range(e2e_steps):
        ts = time.perf_counter()
        out = run_sharded(chains, run_one, rank, world, dist=dist, device=dev, assign=assign)
        e2e_step_ms.append(1e3 * (time.perf_counter() - ts))
    barrier()
    e2e_wall = time.perf_counter() - t1
    if rank == 0:
        for i in mine:
            assert np.array_equal(out[i], res_dev[i]), "e2e and resident runs disagree"
    h2d = sum(cps[i].entry_em.nbytes + chains[i].allele_ref.nbytes + chains[i].n_alleles.nbytes + chains[i].recomb.nbytes +
              chains[i].read_tag.nbytes + chains[i].read_entry_offset.nbytes + chains[i].entry_column.nbytes +
              chains[i].entry_allele.nbytes for i in mine)
    d2h = sum(int(res_dev[i].nbytes) for i in mine)

    # ---- schedule actually run on this rank (from the handles), bytes per direction
    n_fwd = sum(st0[i]["n_forward_steps"] for i in mine)
    n_bwd = sum(st0[i]["n_backward_steps"] for i in mine)
    sched_bytes = float(sum(st0[i]["scheduled_bytes"] for i in mine))
    fwd_bytes = 0.0
    for i in mine:
        p = chains[i]
        u = p.untagged_per_column().astype(np.int64)
        rr = st0[i]["state_columns"] / float(np.sum(2.0 ** u))                  # R^2 of the device layout
        sc = rr * 2.0 ** u
        fwd_bytes += 4.0 * float(2.0 * sc.sum() + sc[:-1].sum())
    bwd_bytes = sched_bytes - fwd_bytes
    my_states = float(sum(st0[i]["state_columns"] for i in mine))
    my_cols = float(sum(chains[i].n_columns for i in mine))

    # ---- max / sum over ranks
    if dist is not None:
        mine_t = torch.tensor([dev_ms / args.steps, my_wall * 1e3 / args.steps, float(clocks.get("sm_mhz") or 0),
                               float(clocks.get("power_w_max") or 0), float(clocks.get("temp_c_max") or 0), my_cols, my_states],
                              dtype=torch.float64, device="cuda")
        allr = [torch.empty_like(mine_t) for _ in range(world)]
        dist.all_gather(allr, mine_t)
        per_rank = [[round(float(x), 1) for x in r.tolist()] for r in allr]
        t = torch.tensor([dev_ms, wall * 1e3, e2e_wall * 1e3, inst_wall * 1e3], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms_max, wall_ms, e2e_ms, inst_ms = [float(x) for x in t.tolist()]
        tot = torch.tensor([fwd_ms, bwd_ms, fwd_bytes, bwd_bytes, float(n_fwd), float(n_bwd), float(launches), float(h2d), float(d2h),
                            sched_bytes], dtype=torch.float64, device="cuda")
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        fwd_ms, bwd_ms, fwd_bytes, bwd_bytes, n_fwd, n_bwd, launches, h2d, d2h, sched_bytes = [float(x) for x in tot.tolist()]
    else:
        wall_ms, e2e_ms, dev_ms_max, inst_ms = wall * 1e3, e2e_wall * 1e3, dev_ms, inst_wall * 1e3
        per_rank = None
    if rank != 0:
        dist.destroy_process_group()
        return

    ms_per_step = wall_ms / args.steps
    value = total_cols * args.steps / (wall_ms / 1e3)
    peak, peak_src = measured_peak_gbs()
    # the column-step kernel family: a forward launch moves 12 B/state (read alpha_{c-1}, read beta_c, write alpha_c), a
    # backward launch 8 B/state (read beta_{c+1}, write beta_c); the engine's `scheduled_bytes` is that sum over every
    # launch of the schedule, recompute launches included.  Summed over ranks: bytes / summed kernel time = the mean
    # per-GPU kernel bandwidth.
    fwd_gbs = fwd_bytes * args.steps / (fwd_ms / 1e3) / 1e9
    bwd_gbs = bwd_bytes * args.steps / (bwd_ms / 1e3) / 1e9
    dominant = "forward" if fwd_ms >= bwd_ms else "backward"
    ach = fwd_gbs if dominant == "forward" else bwd_gbs
    traffic, traffic_note = None, None
    for tname in ("r02_traffic_u15.json", "r01_traffic_u15.json"):
        tpath = os.path.join(ROOT, "profiles", tname)
        if args.workload in ("config2", "genome") and os.path.exists(tpath):
            try:
                cap = json.load(open(tpath))[dominant][0]
                traffic = (cap["dram_read"] + cap["dram_write"]) * 1e9
                alg = (12.0 if dominant == "forward" else 8.0) * (1 << 15) * 88 * 88
                traffic_note = (f"ncu dram__bytes_read+write of one {dominant} launch on a 2^15-block column ({tname}): {traffic:.4g} B "
                                f"against {alg:.4g} algorithmic bytes for that launch ({cap['duration_us']:.0f} us under ncu)")
                break
            except Exception:
                pass
    # the kernel behind the dominant direction (engine.cu picks by panel size: pair_wave_kernel for even panels <= 16,
    # wide_step_kernel above 128 haplotypes, colown_step_kernel for the backward step of 33..128, step_tma_kernel for the
    # forward step)
    if args.workload == "tagged":
        kernel_name = "chain_kernel"
    elif args.haps <= 16 and args.haps % 2 == 0:
        kernel_name = "pair_wave_kernel"
    elif args.haps > 128:
        kernel_name = "wide_step_kernel"
    elif dominant == "backward" and args.haps > 32 and args.haps % 4 == 0 and os.environ.get("GG_COLOWN", "1") != "0":
        kernel_name = "colown_step_kernel"
    else:
        kernel_name = "step_tma_kernel"
    roofline = {
        "bound": "hbm", "kernel": f"step kernel, {dominant} ({kernel_name})", "achieved": ach, "peak": peak, "unit": "GB/s",
        "frac": ach / peak, "box_copy_GBps": box_copy_gbs, "frac_of_box_copy": (ach / box_copy_gbs if box_copy_gbs else None),
        "traffic": traffic, "traffic_note": traffic_note, "peak_source": peak_src,
        "algorithmic_bytes_per_launch": (fwd_bytes / n_fwd) if dominant == "forward" else (bwd_bytes / n_bwd),
        "avg_launch_ms": (fwd_ms / (n_fwd * args.steps)) if dominant == "forward" else (bwd_ms / (n_bwd * args.steps)),
        "forward": {"GBps": fwd_gbs, "frac": fwd_gbs / peak, "launches_per_step": n_fwd, "ms_per_step": fwd_ms / args.steps},
        "backward": {"GBps": bwd_gbs, "frac": bwd_gbs / peak, "launches_per_step": n_bwd, "ms_per_step": bwd_ms / args.steps,
                     "launches_per_column": n_bwd / total_cols},
        # 20 B per state per column, recompute NOT credited, over the slowest rank's device time
        "whole_sweep_algorithmic_GBps": 20.0 * total_states * args.steps / (dev_ms_max / 1e3) / 1e9,
        "whole_sweep_algorithmic_frac": 20.0 * total_states * args.steps / (dev_ms_max / 1e3) / 1e9 / (peak * world),
        "whole_sweep_scheduled_GBps": sched_bytes * args.steps / (dev_ms_max / 1e3) / 1e9,
    }
    if full_length is not None:
        # what the measured per-launch speeds mean for the whole chromosome: forward once and backward
        # `backward_bytes_per_column_byte` times over its state-columns
        S_full = full_length["state_columns"]
        t_full = (12.0 * S_full / (fwd_gbs * 1e9)) + (8.0 * full_length["backward_bytes_per_column_byte"] * S_full / (bwd_gbs * 1e9))
        full_length.update({
            "this_run_backward_bytes_per_column_byte": (bwd_bytes / 8.0) / (fwd_bytes / 12.0),
            "projected_seconds_one_gpu": t_full, "projected_variants_per_s": FULL_LENGTH_COLUMNS / t_full,
            "projected_whole_sweep_algorithmic_frac": 20.0 * S_full / t_full / 1e9 / peak,
            "note": "projection = measured forward / backward kernel GB/s of this run applied to the 516k-column schedule "
                    "(gg_schedule_dry_run on the synthetic coverage profile, this device's default budget)"})
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": config,
        "hmm_columns_per_s": value, "state_columns_per_s": total_states * args.steps / (wall_ms / 1e3),
        "device_ms_per_step": dev_ms_max / args.steps,
        "value_with_per_launch_events": total_cols * args.steps / (inst_ms / 1e3),
        "value_without_launch_events": value,
        "value_note": "timed region = K steps of gg_hmm_sweep over the resident chains; the per-kernel times of `roofline` come "
                      "from K more steps of the same work with a CUDA event pair around every launch (value_with_per_launch_events)",
        "roofline": roofline,
        "full_length": full_length,
        "cpu_baseline": cpu_baseline,
        "e2e": {"value": total_cols * e2e_steps / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "steps": e2e_steps, "step_ms": [round(x, 2) for x in e2e_step_ms],
                "path": "per chain gg_hmm_run (plan + H2D + sweep + D2H)" + (", then gather of posteriors to rank 0" if world > 1 else "")},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "lpt": lpt,
        "per_rank": ({"columns": ["device_ms_per_step", "busy_wall_ms_per_step", "sm_mhz", "power_w_max", "temp_c_max", "variant_sites",
                                  "state_columns"], "rows": per_rank} if per_rank else None),
        "checkpoint_levels": max(st0[i]["checkpoint_levels"] for i in mine),
        "state_bytes_reserved": max(st0[i]["state_bytes_reserved"] for i in mine),
        "max_untagged": max(st0[i]["max_untagged"] for i in mine),
    }
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
