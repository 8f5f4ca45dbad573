#!/usr/bin/env python
"""Benchmark of the GenotypeHMM forward-backward (BASELINE.json metric: variants genotyped/s = HMM
columns/s; achieved HBM GB/s).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU GenotypeHMM on the host cores

Workload (BASELINE.json configs[1]): chr20-scale problem — 88-haplotype panel, 30x HiFi, max-coverage 15,
516k variant sites over 64 Mb — of which one STEP sweeps a window of --columns consecutive variant sites
(default 256; the full chromosome is 2016 such windows of identical shape and would take ~10 min per pass
at the HBM roofline).  A step is one complete forward-backward over the window: every backward column,
every forward column, every genotype posterior; with 256 columns the state does not fit HBM, so the
checkpoint schedule (recompute) is part of the step, as it is at full scale.  Synthetic data, seeded.

N > 1: every rank sweeps its own copy of the window (independent chains of identical size) — there is
no data-path collective ("weak" scaling); the end-to-end arm adds the final gather of posteriors to rank 0.
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
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
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
def cpu_sample_problems(n, sample_cols, sample_cov, n_haps, make=None):
    from giggles_b200.synth import config2_window
    make = make or config2_window
    return [make(sample_cols, seed=1000 + i, max_coverage=sample_cov, n_haps=n_haps) for i in range(n)]


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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config2", choices=["config2", "config1", "tagged", "config4", "config5"],
                    help="config2 (default, the headline): chr20-scale window; the others are side measurements of the "
                         "remaining BASELINE.json shapes")
    ap.add_argument("--columns", type=int, default=None, help="variant sites per window (one step sweeps one window)")
    ap.add_argument("--max-coverage", type=int, default=15)
    ap.add_argument("--haps", type=int, default=88)
    ap.add_argument("--cpu-sample-cols", type=int, default=240)   # ~10 s of CPU work per sample on 16 host cores
    ap.add_argument("--cpu-sample-cov", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    presets = {"config2": (256, 15, 88), "config1": (5000, 15, 10), "tagged": (20000, 15, 88), "config4": (8, 20, 88),
               "config5": (48, 8, 400)}
    d_cols, d_cov, d_haps = presets[args.workload]
    if args.columns is None:
        args.columns = d_cols
    if args.workload != "config2":
        args.max_coverage, args.haps = d_cov, d_haps

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = os.cpu_count() or 1

    from giggles_b200 import synth

    def config2_window(n_columns, seed=2, max_coverage=15, n_haps=88):
        if args.workload == "config1":
            return synth.config1(seed=seed, n_columns=n_columns)
        if args.workload == "tagged":
            return synth.config2_window(n_columns, seed=seed, tags="all", n_haps=n_haps)
        if args.workload == "config5":
            return synth.config5_window(n_columns, seed=seed, n_haps=n_haps, max_coverage=max_coverage)
        if args.workload == "config4":
            return synth.config4_window(n_columns, seed=seed, n_haps=n_haps, max_coverage=max_coverage)
        return synth.config2_window(n_columns, seed=seed, max_coverage=max_coverage, n_haps=n_haps)

    names = {"config2": f"configs[1] chr20-scale window: {args.columns} of 516k variant sites, {args.haps}-haplotype panel, "
                        f"30x HiFi untagged, max-coverage {args.max_coverage} (<= 2^{args.max_coverage} x {args.haps}^2 states per column)",
             "config1": f"configs[0] 1 Mb region: {args.columns} variant sites, 10-haplotype panel, 10x HiFi untagged",
             "tagged": f"configs[1] shape with every read haplotagged (the mode giggles runs in without --keep-untagged): "
                       f"{args.columns} variant sites, {args.haps}-haplotype panel, one {args.haps}^2 block per column",
             "config4": f"configs[3] shape: {args.columns} variant sites of 60x ONT, {args.haps}-haplotype panel, max-coverage "
                        f"{args.max_coverage} (2^{args.max_coverage} bipartition blocks = {(1 << args.max_coverage) * args.haps ** 2 * 4 / 2**30:.1f} GiB per column)",
             "config5": f"configs[4] shape: {args.columns} variant sites, {args.haps}-haplotype panel, 3..16 alleles per site, "
                        f"max-coverage {args.max_coverage}"}
    config = {"workload": names[args.workload],
              "columns_per_step": args.columns, "n_haps": args.haps, "max_coverage": args.max_coverage,
              "parallelism": f"{world} independent chain(s), one per GPU",
              "l2_policy": ("inputs larger than L2 (each column is up to 1 GB; L2 is 126 MB)" if args.workload == "config2"
                            else "side measurement: columns may fit L2; not the headline number")}

    # ---------------------------------------------------------------- reference arm (CPU, rank 0 only)
    if args.impl == "reference":
        if rank != 0:
            return
        p_work = config2_window(args.columns, seed=2, max_coverage=args.max_coverage, n_haps=args.haps)
        states_per_col = p_work.state_columns() / p_work.n_columns
        probs = cpu_sample_problems(cores, args.cpu_sample_cols, args.cpu_sample_cov, args.haps, config2_window)
        vals = []
        for i in range(args.warmup + args.steps):
            sc_s, _, dt, kind = run_cpu_sample(probs, cores)
            if i >= args.warmup:
                vals.append((sc_s, dt))
        sc_s = sum(v[0] for v in vals) / len(vals)
        ms = 1e3 * sum(v[1] for v in vals) / len(vals)
        value = sc_s / states_per_col
        sample = (f"{cores} chains x {args.cpu_sample_cols} columns, R={args.haps}, max-coverage {args.cpu_sample_cov} "
                  f"({sum(p.state_columns() for p in probs):.3g} state-columns per step); measured {sc_s:.4g} state-columns/s, "
                  f"converted to the workload's {states_per_col:.4g} states per column")
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f80 (x87 long double)", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "state_columns_per_s": sc_s}))
        return

    # ---------------------------------------------------------------- our arm
    from giggles_b200 import capi
    if capi.lib().gg_device_count() < 1:
        raise SystemExit("bench.py: no CUDA device — the GenotypeHMM path has no CPU fallback")
    dist = None
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
            p_tmp = config2_window(args.columns, seed=2, max_coverage=args.max_coverage, n_haps=args.haps)
            spc = p_tmp.state_columns() / p_tmp.n_columns
            probs = cpu_sample_problems(cores, args.cpu_sample_cols, args.cpu_sample_cov, args.haps, config2_window)
            sc_s, _, dt, kind = run_cpu_sample(probs, cores)
            cpu_baseline = {"value": sc_s / spc, "unit": UNIT, "cores": cores, "kind": kind,
                            "sample": f"{cores} chains x {args.cpu_sample_cols} columns, R={args.haps}, max-coverage "
                                      f"{args.cpu_sample_cov}: {dt:.1f} s of CPU work, {sc_s:.4g} state-columns/s, converted "
                                      f"to the workload's {spc:.4g} states per column",
                            "state_columns_per_s": sc_s}
        except Exception as e:  # the baseline must never take the GPU number down with it
            cpu_baseline = {"value": None, "unit": UNIT, "cores": cores, "kind": "unavailable", "sample": str(e)}

    # what this device's HBM does today, measured the way the roofline denominator was (context for `roofline.frac`)
    box_copy_gbs = capi.lib().gg_device_copy_gbs(local_rank, 2 << 30, 5)
    problem = config2_window(args.columns, seed=2, max_coverage=args.max_coverage, n_haps=args.haps)   # same shape AND size on every rank (weak scaling)
    cp = capi.CProblem(problem)
    h = capi.Hmm(cp, device=local_rank)
    st0 = h.stats()
    for _ in range(args.warmup):
        h.sweep()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    dev_ms, fwd_ms, bwd_ms, emit_ms = [], 0.0, 0.0, 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        h.sweep(True)              # synchronous: returns after the stream drained; CUDA events inside
        s = h.stats()
        dev_ms.append(s["last_sweep_ms"])
        fwd_ms += s["last_fwd_ms"]; bwd_ms += s["last_bwd_ms"]; emit_ms += s["last_emit_ms"]
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    st = h.stats()
    launches_per_step = st["n_launches"]
    my_ms = sum(dev_ms)             # device time of exactly K steps on this rank
    res_dev = h.fetch().values
    h.close()

    # ---- end to end through the C ABI with host buffers: plan + H2D + sweep + D2H (+ gather for N > 1)
    e2e_steps = max(1, args.steps)
    def gather_posteriors(values):
        if dist is None:
            return
        import torch
        buf = torch.from_numpy(values).cuda()
        out = [torch.empty_like(buf) for _ in range(world)] if rank == 0 else None
        dist.gather(buf, out, dst=0)
        torch.cuda.synchronize()

    for _ in range(2):
        gather_posteriors(capi.hmm_run(cp, device=local_rank).values)   # warm (arena reuse, NCCL gather channel)
    barrier()
    e2e_step_ms = []
    t1 = time.perf_counter()
    for _ in range(e2e_steps):
        ts = time.perf_counter()
        r = capi.hmm_run(cp, device=local_rank)
        gather_posteriors(r.values)
        e2e_step_ms.append(1e3 * (time.perf_counter() - ts))
    barrier()
    e2e_wall = time.perf_counter() - t1
    assert np.array_equal(r.values, res_dev), "e2e and resident runs disagree"
    h2d = (cp.entry_em.nbytes + problem.allele_ref.nbytes + problem.n_alleles.nbytes + problem.recomb.nbytes +
           problem.read_tag.nbytes + problem.read_entry_offset.nbytes + problem.entry_column.nbytes + problem.entry_allele.nbytes)
    d2h = int(r.values.nbytes)

    # ---- max over ranks
    if dist is not None:
        import torch
        mine = torch.tensor([my_ms / args.steps, float(clocks.get("sm_mhz") or 0), float(clocks.get("power_w_max") or 0),
                             float(clocks.get("temp_c_max") or 0)], dtype=torch.float64, device="cuda")
        allr = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = [[round(float(x), 1) for x in r.tolist()] for r in allr]
        t = torch.tensor([my_ms, wall * 1e3, e2e_wall * 1e3], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        my_ms, wall_ms, e2e_ms = [float(x) for x in t.tolist()]
        tot = torch.tensor([float(problem.n_columns), float(st0["state_columns"])], dtype=torch.float64, device="cuda")
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        total_cols, total_states = [float(x) for x in tot.tolist()]
    else:
        wall_ms, e2e_ms = wall * 1e3, e2e_wall * 1e3
        per_rank = None
        total_cols, total_states = float(problem.n_columns), float(st0["state_columns"])
    if rank != 0:
        dist.destroy_process_group()
        return

    ms_per_step = wall_ms / args.steps
    value = total_cols * args.steps / (wall_ms / 1e3)
    peak, peak_src = measured_peak_gbs()
    S = st0["state_columns"]
    # dominant kernel: the column-step kernel family; the forward step moves 12 B/state (read alpha_{c-1}, read
    # beta_c, write alpha_c), the backward step 8 B/state (read beta_c, write beta_{c-1})
    n_fwd, n_bwd = st0["n_forward_steps"], st0["n_backward_steps"]
    # bytes per direction, exactly: a forward launch reads alpha_{c-1} and beta_c and writes alpha_c, a backward launch
    # reads beta_{c+1} and writes beta_c; the engine's `scheduled_bytes` is that sum over every launch of the schedule
    # (recompute launches of the checkpoint schedule included), so backward = scheduled - forward
    u = problem.untagged_per_column().astype(np.int64)
    rr = S / float(np.sum(2.0 ** u))                                  # R^2 of the device layout
    sc = rr * 2.0 ** u
    fwd_bytes = 4.0 * float(2.0 * sc.sum() + sc[:-1].sum())
    bwd_bytes = float(st0["scheduled_bytes"]) - fwd_bytes
    fwd_gbs = fwd_bytes * args.steps / (fwd_ms / 1e3) / 1e9
    bwd_gbs = bwd_bytes * args.steps / (bwd_ms / 1e3) / 1e9
    dominant = "forward" if fwd_ms >= bwd_ms else "backward"
    ach = fwd_gbs if dominant == "forward" else bwd_gbs
    # DRAM traffic of the dominant kernel from the committed `ncu --set full` capture (profiles/, one launch of a
    # u = 15 column of this workload: 2^15 x 88^2 states); null for the side workloads
    traffic, traffic_note = None, None
    tpath = os.path.join(ROOT, "profiles", "r01_traffic_u15.json")
    if args.workload == "config2" and os.path.exists(tpath):
        try:
            cap = json.load(open(tpath))[dominant][0]
            traffic = (cap["dram_read"] + cap["dram_write"]) * 1e9
            alg = (12.0 if dominant == "forward" else 8.0) * (1 << 15) * 88 * 88
            traffic_note = (f"ncu dram__bytes_read+write of one {dominant} launch on a 2^15-block column: {traffic:.4g} B "
                            f"against {alg:.4g} algorithmic bytes for that launch ({cap['duration_us']:.0f} us under ncu)")
        except Exception:
            pass
    roofline = {
        "bound": "hbm", "kernel": f"step kernel, {dominant} (" + ("chain_kernel" if args.workload == "tagged" else "step_tma_kernel") + ")", "achieved": ach, "peak": peak, "unit": "GB/s",
        "frac": ach / peak, "box_copy_GBps": box_copy_gbs, "frac_of_box_copy": (ach / box_copy_gbs if box_copy_gbs else None),
        "traffic": traffic, "traffic_note": traffic_note, "peak_source": peak_src,
        "algorithmic_bytes_per_launch": (fwd_bytes / n_fwd) if dominant == "forward" else (bwd_bytes / n_bwd),
        "avg_launch_ms": (fwd_ms / (n_fwd * args.steps)) if dominant == "forward" else (bwd_ms / (n_bwd * args.steps)),
        "forward": {"GBps": fwd_gbs, "frac": fwd_gbs / peak, "launches_per_step": n_fwd, "ms_per_step": fwd_ms / args.steps},
        "backward": {"GBps": bwd_gbs, "frac": bwd_gbs / peak, "launches_per_step": n_bwd, "ms_per_step": bwd_ms / args.steps},
        "whole_sweep_algorithmic_GBps": 20.0 * total_states * args.steps / (my_ms / 1e3) / 1e9,
        "whole_sweep_scheduled_GBps": st0["scheduled_bytes"] * world * args.steps / (my_ms / 1e3) / 1e9,
    }
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": config,
        "hmm_columns_per_s": value, "state_columns_per_s": total_states * args.steps / (wall_ms / 1e3),
        "device_ms_per_step": my_ms / args.steps,
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
        "e2e": {"value": total_cols * e2e_steps / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": d2h, "steps": e2e_steps, "step_ms": [round(x, 2) for x in e2e_step_ms], "path": "gg_hmm_run (plan + H2D + sweep + D2H)" +
                (" + gather of posteriors to rank 0" if world > 1 else "")},
        "gpu_launches": int(launches_per_step * args.steps),
        "clocks": clocks,
        "per_rank": ({"columns": ["device_ms_per_step", "sm_mhz", "power_w_max", "temp_c_max"], "rows": per_rank} if per_rank else None),
        "checkpoint_levels": st0["checkpoint_levels"], "state_bytes_reserved": st0["state_bytes_reserved"],
        "max_untagged": st0["max_untagged"],
    }
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
