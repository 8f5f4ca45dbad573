"""Bit-identical column structure: the product's planner (gg_plan_*) against the reference's own
ColumnIterator / Column classes (oracle/_ref) and against the restatement's compatible lists."""
import ctypes as C

import numpy as np
import pytest

import oracle
from conftest import golden_files, load_golden
from giggles_b200 import capi
from giggles_b200.synth import make_problem

needs_ref = pytest.mark.skipif(not oracle.have_reference(), reason="oracle/_ref not built (needs /root/reference)")


def _check_against_reference(p, max_bits=8):
    act, unt, comp = oracle.reference_structure(p, max_bits)
    pv = capi.PlanView(p)
    assert pv.n_columns == p.n_columns
    for c in range(p.n_columns):
        assert pv.active(c) == act[c]
        assert pv.untagged(c) == unt[c]
        if c >= 1:
            for b, lst in enumerate(comp[c]):
                assert sorted(pv.compatible_prev(c, b)) == sorted(lst)


@needs_ref
@pytest.mark.parametrize("seed", range(6))
def test_planner_matches_reference_columns(seed):
    tags = ["none", "mixed", "all"][seed % 3]
    p = make_problem(50, 5, coverage=7, max_coverage=7, read_len_mean=1500, tags=tags, seed=40 + seed, gap_frac=0.1,
                     window=bool(seed % 2))
    _check_against_reference(p)


@needs_ref
@pytest.mark.parametrize("path", golden_files()[:8], ids=lambda p: p.split("/")[-1][:-4])
def test_planner_matches_reference_on_golden(path):
    p, _ = load_golden(path)
    _check_against_reference(p)


def test_planner_matches_restatement_compatible_lists():
    # works without the reference: the oracle's compatible-list routine follows src/column.cpp:72-122
    p = make_problem(60, 4, coverage=8, max_coverage=8, read_len_mean=1200, tags="mixed", seed=77, window=True)
    pv = capi.PlanView(p)
    L = oracle.oracle_lib()
    for c in range(1, p.n_columns):
        left = np.array(pv.untagged(c - 1), dtype=np.uint32)
        right = np.array(pv.untagged(c), dtype=np.uint32)
        out = np.zeros(max(2, 2 << len(left)), dtype=np.uint32)
        for b in range(1 << len(right)):
            n = L.oracle_compatible_flat(len(left), left.ctypes.data_as(C.c_void_p), len(right),
                                         right.ctypes.data_as(C.c_void_p), b, out.ctypes.data_as(C.c_void_p))
            assert sorted(out[:n].tolist()) == sorted(pv.compatible_prev(c, b))


def test_untagged_counts_match_problem_helper():
    p = make_problem(80, 4, coverage=6, max_coverage=6, read_len_mean=1500, tags="mixed", seed=5)
    pv = capi.PlanView(p)
    u = p.untagged_per_column()
    assert [len(pv.untagged(c)) for c in range(p.n_columns)] == u.tolist()


def test_unsorted_reads_are_rejected_like_the_reference():
    p = make_problem(30, 4, coverage=5, max_coverage=5, read_len_mean=1500, seed=6)
    assert p.n_reads >= 3
    # swap the first and last read
    off = p.read_entry_offset
    order = [p.n_reads - 1] + list(range(1, p.n_reads - 1)) + [0]
    import copy
    q = copy.copy(p)
    ecol, eall, tag, roff, soff, sc = [], [], [], [0], [0], []
    for k in order:
        for e in range(off[k], off[k + 1]):
            ecol.append(p.entry_column[e]); eall.append(p.entry_allele[e])
            sc.extend(p.entry_scores[p.entry_score_offset[e]:p.entry_score_offset[e + 1]]); soff.append(len(sc))
        roff.append(len(ecol)); tag.append(p.read_tag[k])
    q.entry_column = np.array(ecol, np.uint32); q.entry_allele = np.array(eall, np.int32)
    q.read_tag = np.array(tag, np.int8); q.read_entry_offset = np.array(roff, np.uint32)
    q.entry_score_offset = np.array(soff, np.uint64); q.entry_scores = np.array(sc, np.float64)
    with pytest.raises(capi.GGError, match="not sorted") as ei:
        capi.PlanView(q)
    assert ei.value.status == capi.GG_ERR_INVALID


@needs_ref
def test_planner_matches_reference_up_to_13_bits():
    # the masks the kernels decode with (shared / terminating bits, src/column.cpp:72-122) at bipartition counts where
    # a CTA of the pipelined kernel walks many items: every b of every column against the reference's own lists
    p = make_problem(14, 4, region_bp=14 * 124, coverage=30, max_coverage=13, read_len_mean=500, read_len_sigma=0.5,
                     window=True, seed=1203)
    u = p.untagged_per_column()
    assert int(u.max()) >= 12 and len(set(u.tolist())) >= 3
    _check_against_reference(p, max_bits=13)
