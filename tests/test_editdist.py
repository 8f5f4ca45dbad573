"""Batched device edit distance against the reference's edit_distance (giggles/align.pyx:15-107):
committed golden vectors generated from the reference's own Cython module, the C restatement (oracle/edit_oracle.c)
on fresh random batches, and the reference module itself where oracle/_ref/pyref is present.  Integer results:
bit-exact, in the banded mode too (where the reference's value above maxdiff depends on its band bookkeeping)."""
import json
import os

import numpy as np
import pytest

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "editdist_golden.json")


def _golden():
    with open(GOLDEN) as f:
        return json.load(f)


def test_restatement_matches_golden():
    n = 0
    for s, t, d in _golden():
        for md, want in d.items():
            assert oracle.oracle_edit_distance(s, t, int(md)) == want, (s[:30], t[:30], md)
            n += 1
    assert n > 1000


def test_restatement_matches_reference_module_live():
    ed = oracle.reference_edit_distance()
    if ed is None:
        pytest.skip("oracle/_ref/pyref/giggles/align not built (needs /root/reference)")
    rng = np.random.default_rng(3)
    for _ in range(300):
        a = "".join(rng.choice(list("ACGT"), size=int(rng.integers(0, 200))))
        b = "".join(rng.choice(list("ACGT"), size=int(rng.integers(0, 200))))
        md = int(rng.choice([-1, 0, 2, 5, 20, 100]))
        assert oracle.oracle_edit_distance(a, b, md) == ed(a, b, md)


def test_library_exports_the_entry_point():
    from giggles_b200 import capi
    assert hasattr(capi.lib(), "gg_edit_distance_batch")


def test_no_device_is_an_error_not_a_fallback():
    from giggles_b200 import align, capi
    if capi.lib().gg_device_count() > 0:
        pytest.skip("a device is present")
    with pytest.raises(capi.GGError) as ei:
        align.edit_distance("ACGT", "AGGT")
    assert ei.value.status == capi.GG_ERR_NO_DEVICE


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_device_matches_golden(have_gpu):
    if not have_gpu:
        pytest.fail("no CUDA device: the edit distance has no CPU fallback")
    from giggles_b200 import align
    g = _golden()
    for md in (-1, 0, 1, 3, 10, 50):
        sel = [(s, t, d[str(md)]) for s, t, d in g if str(md) in d]
        got = align.edit_distance_batch([(s, t) for s, t, _ in sel], md, device=0)
        want = np.array([w for _, _, w in sel], dtype=np.int32)
        bad = np.nonzero(got != want)[0]
        assert bad.size == 0, (md, [(sel[i][0][:20], sel[i][1][:20], int(got[i]), int(want[i])) for i in bad[:5]])
    # mixed maxdiff in one batch, and the single-pair call with the reference's signature
    sel = [(s, t, int(md), w) for s, t, d in g[:200] for md, w in d.items()]
    got = align.edit_distance_batch([(s, t) for s, t, _, _ in sel], [m for _, _, m, _ in sel], device=0)
    assert got.tolist() == [w for _, _, _, w in sel]
    assert align.edit_distance("kitten", "sitting") == 3 and align.edit_distance(b"kitten", b"sitting", 1) == 2


@pytest.mark.gpu
def test_device_matches_restatement_on_random_batches(have_gpu):
    if not have_gpu:
        pytest.fail("no CUDA device: the edit distance has no CPU fallback")
    from giggles_b200 import align
    rng = np.random.default_rng(11)
    pairs, mds = [], []
    for k in range(1500):
        n = int(rng.choice([0, 1, 5, 40, 64, 65, 200, 700, 2100, 3000], p=[.02, .03, .2, .25, .05, .05, .2, .1, .05, .05]))
        a = "".join(rng.choice(list("ACGTN"), size=n, p=[.24, .24, .24, .24, .04]))
        if rng.random() < 0.7:      # a mutated copy (the realistic case: query vs the allele it came from)
            b = list(a)
            for _ in range(int(rng.integers(0, max(2, n // 15)))):
                pos = int(rng.integers(0, len(b) + 1))
                r = rng.random()
                if r < 0.34 and b:
                    del b[min(pos, len(b) - 1)]
                elif r < 0.67:
                    b.insert(pos, str(rng.choice(list("ACGT"))))
                elif b:
                    b[min(pos, len(b) - 1)] = str(rng.choice(list("ACGT")))
            b = "".join(b)
        else:
            b = "".join(rng.choice(list("ACGT"), size=int(rng.integers(0, n + 50))))
        pairs.append((a, b))
        mds.append(int(rng.choice([-1, -1, -1, 0, 2, 7, 30])))
    got = align.edit_distance_batch(pairs, mds, device=0)
    want = [oracle.oracle_edit_distance(a, b, m) for (a, b), m in zip(pairs, mds)]
    assert got.tolist() == want
    assert align.edit_distance_batch([], -1, device=0).size == 0
