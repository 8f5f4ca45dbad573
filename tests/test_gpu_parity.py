"""Parity of the CUDA path (through the C ABI) with the reference: golden vectors, the CPU oracle
on seeded problems, and size-independent properties at full column sizes."""
import numpy as np
import pytest

import oracle
from conftest import assert_posteriors_close, golden_files, load_golden, py_determine_genotype, py_gq
from giggles_b200 import capi
from giggles_b200.synth import config2_window, make_problem

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _need_device(have_gpu):
    if not have_gpu:
        pytest.fail("no CUDA device: the GenotypeHMM path has no CPU fallback")


@pytest.mark.parametrize("path", golden_files(), ids=lambda p: p.split("/")[-1][:-4])
def test_golden_fp32(path):
    p, want = load_golden(path)
    r = capi.hmm_run(p, device=0)
    assert_posteriors_close(r.values, want)
    # same called GT, GQ within +-1 (reference giggles/cli/genotype.py:82-95, giggles/vcf.py:850-872)
    calls = r.call(0.0)
    off = p.gl_offsets()
    for c in range(p.n_columns):
        l = want[off[c]:off[c + 1]].tolist()
        idx = py_determine_genotype(l)
        srt = sorted(l)
        if len(srt) >= 2 and abs(srt[-1] - srt[-2]) < 1e-5:
            continue        # near-tie in the reference itself (SURVEY §7 hard part 4): GT undefined at this tolerance
        assert calls.gt_index[c] == idx
        if idx >= 0 and py_gq(l, idx) < 60:       # above ~60 the GQ is dominated by fp32 rounding of 1-p
            assert abs(int(calls.gq[c]) - py_gq(l, idx)) <= 1


@pytest.mark.parametrize("path", golden_files(), ids=lambda p: p.split("/")[-1][:-4])
def test_golden_fp64(path):
    p, want = load_golden(path)
    got = capi.hmm_run(p, device=0, precision=1).values
    assert np.abs(got - want).max() <= 1e-11


CASES = [
    dict(n_columns=40, n_haps=6, coverage=5, max_coverage=6, read_len_mean=2000, tags="none"),
    dict(n_columns=40, n_haps=5, coverage=5, max_coverage=6, read_len_mean=2000, tags="mixed"),
    dict(n_columns=40, n_haps=8, coverage=6, max_coverage=6, read_len_mean=2000, tags="all"),
    dict(n_columns=60, n_haps=10, coverage=8, max_coverage=8, read_len_mean=2500, multiallelic_frac=0.4, max_alleles=6),
    dict(n_columns=30, n_haps=7, coverage=6, max_coverage=7, read_len_mean=2500, unknown_allele_frac=0.1, missing_frac=0.05),
    dict(n_columns=24, n_haps=88, coverage=4, max_coverage=4, read_len_mean=2000, window=True),
    dict(n_columns=16, n_haps=88, coverage=8, max_coverage=6, read_len_mean=3000, tags="mixed", window=True, multiallelic_frac=0.3),
    dict(n_columns=30, n_haps=88, coverage=10, max_coverage=7, read_len_mean=600, missing_frac=0.05),
    dict(n_columns=10, n_haps=36, coverage=6, max_coverage=5, read_len_mean=1500, window=True),
    dict(n_columns=8, n_haps=200, coverage=3, max_coverage=2, read_len_mean=1500, multiallelic_frac=1.0, max_alleles=16, window=True),
    dict(n_columns=6, n_haps=400, coverage=3, max_coverage=1, read_len_mean=1500, multiallelic_frac=1.0, max_alleles=16, window=True),
    dict(n_columns=20, n_haps=33, coverage=5, max_coverage=4, read_len_mean=1500),
    dict(n_columns=40, n_haps=12, coverage=10, max_coverage=9, read_len_mean=1200, gap_frac=0.2),
    dict(n_columns=30, n_haps=64, coverage=8, max_coverage=6, read_len_mean=900, multiallelic_frac=0.5, max_alleles=16),
    dict(n_columns=30, n_haps=128, coverage=6, max_coverage=4, read_len_mean=900, tags="mixed"),
    dict(n_columns=12, n_haps=4, coverage=3, max_coverage=3, read_len_mean=1500),
    dict(n_columns=200, n_haps=10, region_bp=40000, coverage=10, max_coverage=10),
    dict(n_columns=30, n_haps=94, coverage=10, max_coverage=6, read_len_mean=800, multiallelic_frac=0.3, max_alleles=8),   # R % 4 == 2
    dict(n_columns=12, n_haps=130, coverage=8, max_coverage=4, read_len_mean=500, multiallelic_frac=0.5),
]


@pytest.mark.parametrize("i", range(len(CASES)))
@pytest.mark.parametrize("variant", [0, 1], ids=["auto", "generic"])
def test_seeded_problems_against_oracle(i, variant):
    p = make_problem(seed=500 + i, **CASES[i])
    want = oracle.oracle_run(p)
    got = capi.hmm_run(p, device=0, kernel_variant=variant).values
    assert_posteriors_close(got, want)


@pytest.mark.parametrize("i", [0, 3, 5, 11, 13])
def test_fp64_mode_is_tight(i):
    p = make_problem(seed=500 + i, **CASES[i])
    assert np.abs(capi.hmm_run(p, device=0, precision=1).values - oracle.oracle_run(p)).max() <= 1e-11


def test_randomized_campaign_all_kernel_families():
    # seeded random shapes (panel size, coverage, tags, alleles, gaps, missing panel entries) through every kernel
    # family: 0 = auto, 1 = generic, 2 = pipelined/one CTA per SM, 3 = no 16-warp variant, 4 = no chain kernel,
    # 5 = no warp-per-block kernel, 7 = small panels column by column (no wave runs)
    rng = np.random.default_rng(2024)
    done = 0
    while done < 48:
        R = int(rng.choice([4, 6, 9, 10, 12, 16, 24, 33, 34, 36, 64, 66, 88, 94, 128, 130, 132, 200, 254, 256, 400]))
        max_cov = int(rng.integers(0, 8))
        while (1 << max_cov) * R * R > (2e6 if R <= 128 else 4e6) and max_cov > 0:
            max_cov -= 1
        kw = dict(n_columns=int(rng.integers(2, 30 if R <= 128 else 10)), n_haps=R, coverage=float(rng.uniform(0.5, 12)),
                  max_coverage=max(max_cov, 1), read_len_mean=float(rng.choice([400, 800, 1500, 3000])),
                  tags="all" if max_cov == 0 else str(rng.choice(["none", "mixed", "all"])),
                  multiallelic_frac=float(rng.choice([0.0, 0.1, 0.5, 1.0])), max_alleles=int(rng.choice([3, 4, 8, 16])),
                  missing_frac=float(rng.choice([0.0, 0.005, 0.05])), gap_frac=float(rng.choice([0.0, 0.02, 0.3])),
                  unknown_allele_frac=float(rng.choice([0.0, 0.1])), window=bool(rng.integers(2)),
                  seed=int(rng.integers(1 << 30)))
        p = make_problem(**kw)
        want = oracle.oracle_run(p)
        if np.isnan(want).any():
            continue
        for kv in (0, 1, 2, 3, 4, 5, 7):
            assert_posteriors_close(capi.hmm_run(p, device=0, kernel_variant=kv).values, want)
        done += 1


def test_all_tagged_chain_matches_per_column_kernels():
    # the regime of real runs (every read haplotagged, one block per column): chain kernel vs per-column launches
    p = make_problem(400, 88, region_bp=60000, coverage=12, max_coverage=10, read_len_mean=3000, tags="all", seed=77,
                     multiallelic_frac=0.2, max_alleles=6)
    want = capi.hmm_run(p, device=0, precision=1).values
    h = capi.Hmm(p, device=0)
    h.sweep()
    chain = h.fetch().values
    assert h.stats()["n_launches"] < 20            # a handful of launches for 400 columns
    h.close()
    assert_posteriors_close(chain, want)
    assert_posteriors_close(capi.hmm_run(p, device=0, kernel_variant=4).values, want)


def _underflow_problem(n=5, second=-1.0):
    # n reads that all contradict the panel at 1e-13 each: every state of column 0 has emission 1e-65, far below
    # fp32 range; the reference (80-bit) still gets the answer
    from giggles_b200.problem import Problem
    return Problem(positions=[100, 200], n_alleles=[2, 2], allele_ref=[[1, 1, 1, 1], [0, 1, 0, 1]], recomb=[0.01],
                   read_tag=[-1] * n, read_entry_offset=np.arange(0, 2 * n + 1, 2), entry_column=[0, 1] * n,
                   entry_allele=[0, 0] * n, entry_score_offset=np.arange(0, 4 * n + 1, 2),
                   entry_scores=[0.0, -30.0, 0.0, second] * n)


def test_fp32_stays_in_range_when_every_state_underflows():
    # the per-column emission shift (emission_kernel: best state of every column ~1, exact powers of two) keeps the
    # fp32 path in range on its own: no fp64 rerun, tolerance of the fp32 path
    before = capi.lib().gg_fp64_rerun_count()
    for p in (_underflow_problem(), _underflow_problem(12, -30.0)):
        want = oracle.oracle_run(p)
        assert np.isfinite(want).all()
        h = capi.Hmm(p, device=0)              # the staged API has no guard at all
        h.sweep()
        raw32 = h.fetch().values
        h.close()
        assert np.isfinite(raw32).all()
        assert_posteriors_close(raw32, want)
        r = capi.hmm_run(p, device=0)
        assert not r.reran_fp64 and np.array_equal(r.values, raw32)
        rb = capi.hmm_run_batch([p, p], device=0)
        assert all(not x.reran_fp64 and np.array_equal(x.values, raw32) for x in rb)
    assert capi.lib().gg_fp64_rerun_count() == before


def test_batch_of_chains_matches_single_runs():
    # gg_hmm_run_batch: chromosomes back to back, planning of chain i+1 overlapped with the sweep of chain i
    probs = [make_problem(30 + 11 * i, [12, 88, 10, 36][i % 4], coverage=5, max_coverage=4 + i % 3, read_len_mean=1500,
                          tags=["none", "all", "mixed"][i % 3], seed=300 + i, window=bool(i % 2)) for i in range(6)]
    singles = [capi.hmm_run(p, device=0).values for p in probs]
    batch = capi.hmm_run_batch(probs, device=0)
    assert len(batch) == len(probs)
    for a, b in zip(singles, batch):
        assert np.array_equal(a, b.values)
    assert capi.hmm_run_batch([], device=0) == []
    # a malformed chain in the middle: its status and number come back, whatever the other chains did meanwhile
    import copy
    bad = copy.copy(probs[2])
    bad.n_alleles = probs[2].n_alleles.copy()
    bad.n_alleles[1] = 40
    with pytest.raises(capi.GGError) as ei:
        capi.hmm_run_batch(probs[:2] + [bad] + probs[3:], device=0)
    assert ei.value.status == capi.GG_ERR_INVALID and "chain 2" in str(ei.value)
    # more chains than worker threads, sizes far apart (small tagged chains next to windows that take GBs)
    many = [make_problem(20 + (7 * i) % 50, 88, coverage=6, max_coverage=[0, 0, 9, 0][i % 4] or 1, read_len_mean=2500,
                         tags=["all", "all", "none", "all"][i % 4], seed=700 + i, window=True) for i in range(40)]
    res = capi.hmm_run_batch(many, device=0)
    for p, r in zip(many, res):
        assert np.array_equal(capi.hmm_run(p, device=0).values, r.values)


def test_single_column_and_empty_problem():
    p = make_problem(1, 6, coverage=0.0, seed=3)
    assert_posteriors_close(capi.hmm_run(p, device=0).values, oracle.oracle_run(p))
    from giggles_b200.problem import Problem
    e = Problem(positions=[], n_alleles=[], allele_ref=np.zeros((0, 4), np.int32), recomb=[], read_tag=[],
                read_entry_offset=[0], entry_column=[], entry_allele=[], entry_score_offset=[0], entry_scores=[])
    assert capi.hmm_run(e, device=0).values.size == 0


def test_single_entry_reads_on_one_column():
    # reads with exactly one variant: every read starts and ends on the same column
    from giggles_b200.problem import Problem
    n = 5
    p = Problem(positions=[100, 200, 300], n_alleles=[2, 3, 2], allele_ref=[[0, 1, 1, 0], [2, 0, 1, 1], [0, 0, 1, -1]],
                recomb=[0.01, 0.02], read_tag=[-1, 0, -1, 1, -1], read_entry_offset=np.arange(n + 1),
                entry_column=[0, 1, 1, 1, 2], entry_allele=[0, 2, 1, 1, 1],
                entry_score_offset=[0, 2, 5, 8, 11, 13],
                entry_scores=[0, -4, -5, -3, 0, -2, 0, -6, -1, 0, -3, -5, 0])
    assert_posteriors_close(capi.hmm_run(p, device=0).values, oracle.oracle_run(p))


def test_checkpoint_schedule_does_not_change_results():
    p = make_problem(96, 12, coverage=6, max_coverage=6, read_len_mean=2500, seed=5)
    h = capi.Hmm(p, device=0)
    h.sweep()
    base = h.fetch().values
    full = h.stats()["state_bytes_reserved"]
    h.close()
    seen_levels = set()
    for frac in (0.6, 0.35, 0.25):
        h = capi.Hmm(p, device=0, budget=int(full * frac))
        h.sweep()
        st = h.stats()
        seen_levels.add(st["checkpoint_levels"])
        assert st["state_bytes_reserved"] <= full * frac
        assert np.array_equal(h.fetch().values, base)          # bit-identical, not merely close
        h.close()
    assert max(seen_levels) >= 2


@pytest.mark.parametrize("n_haps,slots", [(12, 3.3), (12, 1.3), (88, 3.3), (88, 2.3), (10, 3.3)])
def test_exact_plan_for_few_huge_columns_does_not_change_results(n_haps, slots, monkeypatch):
    monkeypatch.setenv("GG_EXACT_MIN_COL", "0")       # the planner reserves the recursion for columns >= 4 GiB
    # room for alpha x 2 and `slots` of the largest column: the exact planner (nested checkpoints, idle alpha buffer
    # as the sweep's temporary, all-ones last column re-created on demand) takes over — the configs[3] regime in
    # miniature.  Same bits as the everything-fits schedule.
    from giggles_b200 import fullscale as fs
    p = make_problem(40, n_haps, coverage=6, max_coverage=6, read_len_mean=2500, seed=11, window=True)
    cb, fb = fs.column_bytes(p.untagged_per_column(), n_haps, n_alleles=p.n_alleles)
    budget = int((2 + slots) * int(cb.max()) + 4 * int(fb.max()) + 8192)
    dry = capi.schedule_dry_run(cb, fb, budget)
    assert dry["levels"] >= 3 and dry["n_bwd"] > 1.5 * p.n_columns      # deep recompute, several re-created last columns
    base = capi.hmm_run(p, device=0).values
    h = capi.Hmm(p, device=0, budget=budget)
    h.sweep()
    st = h.stats()
    assert st["state_bytes_reserved"] <= budget and st["checkpoint_levels"] >= 3
    assert np.array_equal(h.fetch().values, base)
    h.close()


def test_runs_are_deterministic():
    p = make_problem(40, 88, coverage=6, max_coverage=6, read_len_mean=1500, seed=9, window=True)
    a = capi.hmm_run(p, device=0).values
    b = capi.hmm_run(p, device=0).values
    assert np.array_equal(a, b)


def test_budget_too_small_reports_nomem():
    p = make_problem(64, 12, coverage=6, max_coverage=6, read_len_mean=2500, seed=5)
    with pytest.raises(capi.GGError) as ei:
        capi.Hmm(p, device=0, budget=50_000)
    assert ei.value.status == capi.GG_ERR_NOMEM


# ---- full-size columns (config 2 shape: R = 88, up to 2^12..2^15 bipartitions): properties that need no oracle
def _full_size_problem(max_cov, n_columns=24, seed=2):
    return config2_window(n_columns, seed=seed, max_coverage=max_cov)


@pytest.mark.parametrize("max_cov", [12, 15])
def test_full_size_properties(max_cov):
    p = _full_size_problem(max_cov)
    r32 = capi.hmm_run(p, device=0)
    off = p.gl_offsets()
    v = r32.values
    assert not np.isnan(v).any() and (v >= 0).all()
    sums = np.add.reduceat(v, off[:-1].astype(np.int64))
    assert np.abs(sums - 1.0).max() < 1e-12                      # each column is a distribution
    # fp64 device run of the same problem: the tolerance-matched cross-check at sizes the CPU cannot reach
    r64 = capi.hmm_run(p, device=0, precision=1)
    assert_posteriors_close(v, r64.values)
    # both kernel families agree
    rg = capi.hmm_run(p, device=0, kernel_variant=1)
    assert_posteriors_close(rg.values, r64.values)
    # schedule invariance at full size (column sizes vary 8x inside the window, so how far the budget can shrink
    # depends on the window: take the tightest of a few budgets that still admits a schedule)
    h = capi.Hmm(p, device=0)
    full = h.stats()["state_bytes_reserved"]
    h.close()
    done = False
    for frac in (0.45, 0.6, 0.75, 0.9):
        try:
            h = capi.Hmm(p, device=0, budget=int(full * frac))
        except capi.GGError as e:
            assert e.status == capi.GG_ERR_NOMEM
            continue
        h.sweep()
        assert h.stats()["checkpoint_levels"] >= 2
        assert np.array_equal(h.fetch().values, v)
        h.close()
        done = True
        break
    assert done


def test_config4_shape_2pow20_bipartitions():
    # configs[3]: max-coverage 20 => 2^20 bipartition blocks per column (small panel so that fp64 fits beside it)
    from giggles_b200.synth import config4_window
    p = config4_window(n_columns=6, n_haps=12, max_coverage=20)
    assert int(p.untagged_per_column().max()) == 20
    r32 = capi.hmm_run(p, device=0)
    r64 = capi.hmm_run(p, device=0, precision=1)
    assert_posteriors_close(r32.values, r64.values)
    off = p.gl_offsets()
    sums = np.add.reduceat(r32.values, off[:-1].astype(np.int64))
    assert np.abs(sums - 1.0).max() < 1e-12


def test_config4_shape_downscaled_against_oracle_under_the_exact_plan(monkeypatch):
    # configs[3] down-scaled to what the oracle can walk (R = 12, max-coverage 16: 2^16 blocks per column) with weak
    # per-read evidence (unsaturated posteriors), run the way the real configuration runs: room for alpha x 2 and
    # three backward columns, i.e. under the exact planner with the re-created last column.  Against the oracle, and
    # bit-identical to the everything-fits schedule.
    from giggles_b200 import fullscale as fs
    monkeypatch.setenv("GG_EXACT_MIN_COL", "0")
    p = make_problem(6, 12, region_bp=6 * 124, coverage=60.0, max_coverage=16, read_len_mean=30000.0, read_len_sigma=0.6,
                     tags="none", window=True, seed=4, base_const=1.1, accuracy=0.8)
    assert int(p.untagged_per_column().max()) == 16
    want = oracle.oracle_run(p)
    assert np.mean((want > 1e-4) & (want < 0.999)) > 0.3
    base = capi.hmm_run(p, device=0).values
    assert_posteriors_close(base, want)
    cb, fb = fs.column_bytes(p.untagged_per_column(), 12, n_alleles=p.n_alleles)
    budget = int(5.3 * int(cb.max()) + 4 * int(fb.max()) + 8192)
    dry = capi.schedule_dry_run(cb, fb, budget)
    assert dry["levels"] >= 2 and dry["n_bwd"] > p.n_columns
    h = capi.Hmm(p, device=0, budget=budget)
    h.sweep()
    assert h.stats()["state_bytes_reserved"] <= budget
    assert np.array_equal(h.fetch().values, base)
    h.close()


def test_config5_shape_400_haplotypes_multiallelic():
    # configs[4]: 400-haplotype panel (160,000 ordered pairs), 3..16 alleles at every site
    from giggles_b200.synth import config5_window
    p = config5_window(n_columns=6, n_haps=400, max_coverage=3)
    want = oracle.oracle_run(p)
    assert_posteriors_close(capi.hmm_run(p, device=0).values, want)
    assert np.abs(capi.hmm_run(p, device=0, precision=1).values - want).max() <= 1e-11


def test_config1_shape_full_coverage_slice():
    # configs[0]: R = 10, 10x HiFi; a slice at the config's own density and coverage cap
    from giggles_b200.synth import config1
    p = config1(n_columns=150)
    want = oracle.oracle_run(p)
    r = capi.hmm_run(p, device=0)
    assert_posteriors_close(r.values, want)
    calls = r.call(0.0)
    off = p.gl_offsets()
    for c in range(p.n_columns):
        l = want[off[c]:off[c + 1]].tolist()
        srt = sorted(l)
        if abs(srt[-1] - srt[-2]) >= 1e-5:
            assert calls.gt_index[c] == py_determine_genotype(l)


def test_read_relabelling_symmetry_full_size():
    # swapping the two sides of EVERY read and transposing nothing else must leave genotype posteriors of
    # bi-allelic columns unchanged (the model is symmetric under (b,R1,R2) <-> (~b,R2,R1))
    p = make_problem(16, 88, region_bp=2000, coverage=12, max_coverage=10, read_len_mean=900, tags="mixed", seed=31,
                     multiallelic_frac=0.0, window=True)
    import copy
    q = copy.copy(p)
    t = p.read_tag.copy()
    t[p.read_tag == 0] = 1
    t[p.read_tag == 1] = 0
    q.read_tag = t
    a = capi.hmm_run(p, device=0, precision=1).values
    b = capi.hmm_run(q, device=0, precision=1).values
    assert np.abs(a - b).max() < 1e-10


_L2_KNOB = r"""
import sys, hashlib
sys.path.insert(0, sys.argv[1])
from giggles_b200 import capi
from giggles_b200.synth import config2_window
p = config2_window(6, max_coverage=11)
print(hashlib.sha256(capi.hmm_run(p, device=0).values.tobytes()).hexdigest())
"""


def test_l2_lookahead_is_a_pure_hint():
    """The producer's L2 look-ahead (GG_L2_AHEAD, hmm_tma_kernel.cuh) only moves data towards L2: with it off, on,
    or looking further ahead, the posteriors are bit-identical.  (Read once per process, hence subprocesses.)"""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    digests = set()
    for mode in ("0", "2817", "513", "3"):
        out = subprocess.run([sys.executable, "-c", _L2_KNOB, root], env=dict(os.environ, GG_L2_AHEAD=mode),
                             capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stderr[-2000:]
        digests.add(out.stdout.strip().splitlines()[-1])
    assert len(digests) == 1


def test_concurrent_chains_match_single_runs():
    """capi.hmm_run_many: independent chains swept concurrently from host threads (own stream each, pooled arenas)
    give bit-identical posteriors to one call per chain, in order."""
    probs = [make_problem(n_columns=30 + 7 * i, n_haps=[8, 12, 88, 10, 33, 88][i % 6], coverage=5, max_coverage=[0, 4, 3, 6, 4, 0][i % 6] or 1,
                          read_len_mean=1500, tags=["all", "none", "mixed"][i % 3], seed=900 + i) for i in range(12)]
    single = [capi.hmm_run(p, device=0).values for p in probs]
    for workers in (3, 8):
        many = capi.hmm_run_many(probs, device=0, workers=workers)
        assert len(many) == len(probs)
        for a, b in zip(single, many):
            assert np.array_equal(a, b.values)
