"""The CPU oracle (oracle/hmm_oracle.c) against the reference's golden vectors and, when
oracle/_ref is built, against the compiled reference itself."""
import numpy as np
import pytest

import oracle
from conftest import golden_files, load_golden
from giggles_b200.synth import make_problem


@pytest.mark.parametrize("path", golden_files(), ids=lambda p: p.split("/")[-1][:-4])
def test_oracle_matches_reference_golden(path):
    p, want = load_golden(path)
    got = oracle.oracle_run(p)
    assert np.abs(got - want).max() <= 1e-12


def test_golden_cover_the_edge_cases():
    names = {p.split("/")[-1][:-4] for p in golden_files()}
    assert {"no_reads", "single_column", "all_tagged", "multiallelic", "missing_panel", "gaps_unknown",
            "long_chain_ckpt", "all16_alleles"} <= names
    assert len(names) >= 20


needs_ref = pytest.mark.skipif(not oracle.have_reference(), reason="oracle/_ref not built (needs /root/reference)")


@needs_ref
@pytest.mark.parametrize("seed", range(8))
def test_oracle_matches_compiled_reference(seed):
    tags = ["none", "mixed", "all"][seed % 3]
    p = make_problem(20 + 7 * seed, 4 + seed, coverage=5, max_coverage=5 + seed % 3, read_len_mean=1800, tags=tags,
                     seed=900 + seed, multiallelic_frac=0.3, max_alleles=5, unknown_allele_frac=0.05 * (seed % 2),
                     missing_frac=0.02 * (seed % 4))
    assert np.abs(oracle.oracle_run(p) - oracle.reference_run(p)).max() <= 1e-13


@needs_ref
def test_entry_emission_is_bit_exact():
    from giggles_b200 import capi
    rng = np.random.default_rng(0)
    for n in (2, 3, 7, 16):
        for reg, base in ((10, float(np.e)), (6, 2.0), (3, 10.0)):
            sc = -rng.integers(0, 12, size=n).astype(np.float64)
            ref = oracle.reference_entry_emission(sc, reg, base)
            assert np.array_equal(capi.entry_emission(sc, reg, base), ref)     # product helper: bit-exact
            assert np.array_equal(oracle.oracle_entry_emission(sc, reg, base), ref)
            assert abs(ref.sum() - 1.0) < 1e-15


def test_posteriors_are_distributions():
    p = make_problem(40, 6, coverage=5, max_coverage=6, read_len_mean=2000, seed=3)
    l = oracle.oracle_run(p)
    off = p.gl_offsets()
    for c in range(p.n_columns):
        assert abs(l[off[c]:off[c + 1]].sum() - 1.0) < 1e-12


def test_empty_readset_gives_panel_prior():
    # SURVEY B-4: no reads => posterior is the panel's genotype frequencies
    p = make_problem(5, 4, coverage=0.0, seed=1, missing_frac=0.0, multiallelic_frac=0.0)
    l = oracle.oracle_run(p)
    off = p.gl_offsets()
    for c in range(p.n_columns):
        a = p.allele_ref[c]
        f1 = (a == 1).mean()
        want = np.array([(1 - f1) ** 2, 2 * f1 * (1 - f1), f1 ** 2])
        # the chain couples neighbouring columns only weakly at these recombination costs; compare the sum rule
        assert abs(l[off[c]:off[c + 1]].sum() - 1) < 1e-12
        assert l[off[c]:off[c + 1]].shape == want.shape
