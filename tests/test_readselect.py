"""Read selection: gg_readselection against the reference's readselection()
(giggles/readselect.pyx:231-271).  Host only — no GPU needed.

The bar is set equality with the reference on every case, including the ones where equally scored reads are ranked
by the iteration order of Python sets (see giggles_b200/csrc/pyset.hpp)."""
import os
import random
import subprocess
import sys

import numpy as np
import pytest

from giggles_b200 import capi
from giggles_b200.readselect import FlatRead, FlatVariant, readselection

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "readselect", "cases.npz")
PYREF = os.path.join(ROOT, "oracle", "_ref", "pyref")


def _golden_cases():
    d = np.load(GOLDEN)
    cur = {k: 0 for k in ("off", "pos", "qual", "src", "pref", "sel")}
    for i in range(len(d["max_cov"])):
        c = {}
        for k in cur:
            n = int(d[k + "_len"][i])
            c[k] = d[k][cur[k]:cur[k] + n]
            cur[k] += n
        c["max_cov"], c["bridging"] = int(d["max_cov"][i]), bool(d["bridging"][i])
        c["pref"] = None if c["pref"].tolist() == [-1] else c["pref"].tolist()
        yield i, c


def test_golden_selections_from_the_reference():
    n = ties = 0
    for i, c in _golden_cases():
        got = capi.readselection_flat(c["off"], c["pos"], c["qual"], c["src"], c["max_cov"], c["pref"], c["bridging"])
        assert got.tolist() == c["sel"].tolist(), f"case {i}"
        n += 1
        ties += len(set(c["qual"].tolist())) == 1
    assert n == 160 and ties >= 30     # a good share of the cases is tie-dominated (single quality value)


def test_pyset_model_follows_the_interpreters_set():
    """The order-exact set model against CPython's own set, after every step of random operation scripts."""
    for seed in range(150):
        rnd = random.Random(seed)
        maxv = rnd.choice([20, 100, 1000, 5000, 70000])
        S = [set() for _ in range(8)]
        ops, want = [], []
        for _ in range(150):
            op = rnd.choice([0, 1, 1, 1, 1, 2, 2, 2, 3, 3, 4, 5, 6, 6, 7, 7, 8])
            d = rnd.randrange(8)
            if op == 0:
                S[d] = set(); ops += [0, d]
            elif op == 1:
                v = rnd.randrange(maxv); S[d].add(v); ops += [1, d, v]
            elif op == 2:
                v = rnd.choice(sorted(S[d])) if S[d] and rnd.random() < 0.8 else rnd.randrange(maxv)
                S[d].discard(v); ops += [2, d, v]
            elif op == 3:
                l = [rnd.randrange(maxv) for _ in range(rnd.randrange(0, 40))]
                S[d].update(l); ops += [3, d, len(l)] + l
            elif op == 4:
                s = rnd.randrange(8); S[d].update(S[s]); ops += [4, d, s]
            elif op == 5:
                s = rnd.randrange(8); S[d] = set(S[s]); ops += [5, d, s]
            elif op == 6:
                s = rnd.randrange(8); S[d] -= S[s]; ops += [6, d, s]
            elif op == 7:
                a, b = rnd.randrange(8), rnd.randrange(8); S[d] = S[a].difference(S[b]); ops += [7, d, a, b]
            else:
                n = rnd.randrange(0, maxv); S[d] = set(range(n)); ops += [8, d, n]
            ops += [9, d]
            want += [len(S[d])] + list(S[d])
        assert capi.pyset_replay(ops, len(want) + 8) == want, f"seed {seed}"


def _random_reads(rnd, npos, nreads, maxlen, quals):
    pos = sorted(rnd.sample(range(10, 50 * npos), npos))
    reads = []
    for _ in range(nreads):
        a = rnd.randrange(0, npos - 1)
        span = pos[a:a + rnd.randrange(2, maxlen + 1)]
        if len(span) < 2:
            span = pos[-2:]
        reads.append(FlatRead(rnd.randrange(3), [FlatVariant(p, rnd.choice(quals)) for p in span]))
    reads.sort(key=lambda r: r.variants[0].position)
    return reads


def test_mirror_signature_and_errors():
    rnd = random.Random(3)
    reads = _random_reads(rnd, 60, 80, 6, [10, 20])
    sel = readselection(reads, 4)
    assert isinstance(sel, set) and sel and sel <= set(range(len(reads)))
    assert readselection(reads, 4, preferred_source_ids=None, bridging=True) == sel
    assert readselection(reads, 4, {1}, False) <= set(range(len(reads)))
    assert readselection([], 5) == set()
    with pytest.raises(ValueError, match="at least two variants"):       # giggles/readselect.pyx:249-252
        readselection(reads + [FlatRead(0, [FlatVariant(7, 10)])], 4)


def test_coverage_cap_holds_at_scale():
    """20,000 reads over 50,000 variants, 30x capped at 15: no variant index inside a selected read's span is used
    more than max_cov times (giggles/coverage.py), every read is decided, and the call is deterministic."""
    rnd = random.Random(11)
    npos, max_cov = 50000, 15
    pos = np.sort(np.asarray(rnd.sample(range(10, 200 * npos), npos), dtype=np.int64))
    starts = np.sort(np.asarray([rnd.randrange(0, npos - 2) for _ in range(20000)]))
    off, vp = [0], []
    for a in starts:
        n = max(2, int(rnd.gauss(75, 15)))
        span = pos[a:a + n] if a + n <= npos else pos[-2:]
        vp.extend(span.tolist())
        off.append(len(vp))
    vq = [rnd.choice([20, 30, 40]) for _ in vp]
    src = [0] * len(starts)
    sel = capi.readselection_flat(off, vp, vq, src, max_cov)
    again = capi.readselection_flat(off, vp, vq, src, max_cov)
    assert sel.tolist() == again.tolist() and 0 < sel.size < len(starts)
    cov = np.zeros(npos + 1, dtype=np.int64)
    vp = np.asarray(vp)
    for r in sel.tolist():
        b = np.searchsorted(pos, vp[off[r]])
        e = np.searchsorted(pos, vp[off[r + 1] - 1]) + 1
        cov[b] += 1
        cov[e] -= 1
    cov = np.cumsum(cov)[:-1]
    assert cov.max() <= max_cov
    assert (cov > 0).mean() > 0.99          # and the cap did not starve the region


_LIVE = r"""
import json, random, sys
sys.path.insert(0, sys.argv[1])
from giggles.core import Read, ReadSet, readselection
out = []
for seed in range(int(sys.argv[2])):
    rnd = random.Random(seed)
    npos = rnd.choice([8, 20, 60, 200]); pos = sorted(rnd.sample(range(10, 50 * npos), npos))
    rs = ReadSet(); flat = []
    for k in range(rnd.choice([3, 10, 40, 150])):
        a = rnd.randrange(0, npos - 1); span = pos[a:a + rnd.randrange(2, 9)]
        if len(span) < 2: span = pos[-2:]
        r = Read("r%d" % k, 60, rnd.randrange(3), 0)
        for p in span: r.add_variant(p, 0, [0.0, -5.0], rnd.choice([10, 20]))
        rs.add(r)
    rs.sort()
    mc, pref, br = rnd.choice([1, 2, 5, 15]), rnd.choice([None, [1]]), rnd.choice([True, False])
    sel = sorted(readselection(rs, mc, None if pref is None else set(pref), br))
    out.append(dict(reads=[[r.source_id, [[v.position, v.quality] for v in r]] for r in rs], mc=mc, pref=pref, br=br, sel=sel))
print(json.dumps(out))
"""


@pytest.mark.skipif(not os.path.isdir(os.path.join(PYREF, "giggles")), reason="oracle/_ref/pyref not built")
def test_live_against_the_reference_module():
    """Fresh random cases through the reference's own Cython module (subprocess: it is a different `giggles` package)."""
    import json
    out = subprocess.run([sys.executable, "-c", _LIVE, PYREF, "120"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    for i, c in enumerate(json.loads(out.stdout.splitlines()[-1])):
        reads = [FlatRead(s, [FlatVariant(p, q) for p, q in vs]) for s, vs in c["reads"]]
        got = readselection(reads, c["mc"], None if c["pref"] is None else set(c["pref"]), c["br"])
        assert sorted(got) == c["sel"], f"live case {i}"
