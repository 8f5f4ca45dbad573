"""C-ABI surface: every declared symbol is exported, host helpers match the reference formulae,
and the compute entry points fail loudly without a device (no CPU fallback)."""
import ctypes as C
import math
import os
import re

import numpy as np
import pytest

from conftest import py_determine_genotype, py_gq
from giggles_b200 import capi
from giggles_b200.synth import make_problem

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "giggles_b200.h")).read()
    declared = set(re.findall(r"\b(gg_[a-z0-9_]+)\s*\(", header))
    declared -= {"gg_status"}
    L = capi.lib()
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, missing
    assert declared == set(capi.SYMBOLS)


def test_no_cpu_fallback_without_device(have_gpu):
    if have_gpu:
        pytest.skip("a device is present")
    p = make_problem(10, 4, coverage=3, max_coverage=3, read_len_mean=1500, seed=1)
    with pytest.raises(capi.GGError) as ei:
        capi.hmm_run(p)
    assert ei.value.status == capi.GG_ERR_NO_DEVICE
    with pytest.raises(capi.GGError):
        capi.Hmm(p)


def test_batch_reports_errors(have_gpu):
    good = make_problem(10, 4, coverage=3, max_coverage=3, read_len_mean=1500, seed=1)
    import copy
    bad = copy.copy(good)
    bad.n_alleles = good.n_alleles.copy()
    bad.n_alleles[2] = 40
    with pytest.raises(capi.GGError) as ei:
        capi.hmm_run_batch([bad, good])
    assert ei.value.status == capi.GG_ERR_INVALID          # a malformed first chain is reported with or without a GPU
    if not have_gpu:
        with pytest.raises(capi.GGError) as ei:
            capi.hmm_run_batch([good, good])
        assert ei.value.status == capi.GG_ERR_NO_DEVICE


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "giggles_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".pyx", ".cpp", ".cu", ".cuh", ".hpp", ".h")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "import oracle" not in src and "from oracle" not in src and "hmm_oracle" not in src, f


def test_binomial_coefficient():
    L = capi.lib()
    for n in range(0, 20):
        for k in range(-1, 22):
            want = math.comb(n, k) if 0 <= k <= n else 0
            assert L.gg_binomial_coefficient(n, k) == want
    assert L.gg_binomial_coefficient(-1, 0) == 0


def test_transition_scalars_follow_the_reference_formula():
    # reference src/transitionprobabilitycomputer.cpp:10-15: fp32 cost, fp64 arithmetic
    for cost, R in ((0.00624, 88), (0.5, 10), (0.0, 4), (12.5, 400)):
        pr, qr = capi.transition_scalars(cost, R)
        r = float(np.float32(cost))
        s = float(R)
        assert pr == (1.0 - math.exp(-(r / s))) / s
        assert qr == math.exp(-(r / s)) + pr
        assert abs((qr - pr) + R * pr - 1.0) < 1e-12   # a row of the per-haplotype transition sums to 1


def test_entry_emission_formula():
    # reference src/entry.cpp:50-64
    sc = np.array([0.0, -3.0, -5.0])
    em = capi.entry_emission(sc, 10, float(np.e))
    p = np.array([1e-10 + 1.0 / sum(math.e ** (x - y) for x in sc) for y in sc])
    assert np.allclose(em, p / p.sum(), rtol=1e-14)
    assert abs(em.sum() - 1) < 1e-15 and em.argmax() == 0


def test_result_call_matches_python_rules():
    from giggles_b200.csrc_testhooks import make_result   # thin ctypes helper over gg_result_* (test only)
    rng = np.random.default_rng(4)
    cols = []
    for n in (2, 2, 3, 4, 2):
        ng = n * (n + 1) // 2
        x = rng.random(ng) ** 4
        cols.append(x / x.sum())
    cols.append(np.array([0.5, 0.5, 0.0]))                 # tie -> no call
    cols.append(np.array([1.0, 0.0, 0.0]))                 # certain -> GQ 10000, GL floor
    cols.append(np.array([0.25, 0.5, 0.25]))
    r = make_result(cols)
    for thr in (0.0, 0.2):
        calls = r.call(thr)
        for c, l in enumerate(cols):
            idx = py_determine_genotype(list(l), thr)
            assert calls.gt_index[c] == idx
            assert calls.gq[c] == py_gq(list(l), idx)
            gl = calls.gl_log10[int(r.offsets[c]):int(r.offsets[c + 1])]
            want = [max(math.log10(x), -1000) if x > 0 else -1000 for x in l]
            assert np.allclose(gl, want, rtol=0, atol=1e-12)


def test_problem_validation_errors():
    p = make_problem(12, 4, coverage=3, max_coverage=3, read_len_mean=1500, seed=2)
    import copy
    q = copy.copy(p)
    q.allele_ref = p.allele_ref.copy()
    q.allele_ref[0, 0] = 7      # >= n_alleles
    with pytest.raises(capi.GGError, match="allele_ref"):
        capi.PlanView(q)
    q = copy.copy(p)
    q.n_alleles = p.n_alleles.copy()
    q.n_alleles[3] = 17
    with pytest.raises(capi.GGError, match="n_alleles"):
        capi.PlanView(q)


def test_too_many_untagged_reads_is_reported():
    # 26 untagged reads over two columns: beyond GG_MAX_UNTAGGED
    n = 26
    from giggles_b200.problem import Problem
    p = Problem(positions=[10, 20], n_alleles=[2, 2], allele_ref=[[0, 1], [0, 1]], recomb=[0.01],
                read_tag=[-1] * n, read_entry_offset=np.arange(0, 2 * n + 1, 2), entry_column=[0, 1] * n,
                entry_allele=[0, 0] * n, entry_score_offset=np.arange(0, 4 * n + 1, 2), entry_scores=[0.0, -3.0] * (2 * n))
    with pytest.raises(capi.GGError) as ei:
        capi.PlanView(p)
    assert ei.value.status == capi.GG_ERR_UNSUPPORTED
