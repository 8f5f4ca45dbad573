"""Generates tests/golden/readselect/cases.npz from the UNMODIFIED reference (giggles.core.readselection, built by
`oracle/build_ref.sh --python` into oracle/_ref/pyref).  Run in the build container:

    python tests/golden/make_readselect_golden.py

Each case is a random ReadSet (flat arrays), the call's parameters, and the index set the reference selected.
Many cases are built to be full of score ties (one quality value, equal read lengths), which the reference breaks by
container iteration order — the part a port is most likely to get wrong."""
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref", "pyref"))
from giggles.core import Read, ReadSet, readselection  # noqa: E402  (the reference)


def make_case(seed):
    rnd = random.Random(seed)
    npos = rnd.choice([8, 20, 60, 200, 500])
    pos = sorted(rnd.sample(range(10, 50 * npos), npos))
    nreads = rnd.choice([3, 10, 40, 150, 400])
    maxlen = rnd.choice([3, 6, 12, 30])
    quals = rnd.choice([[10], [10, 20], list(range(5, 40))])
    gap = rnd.choice([0.0, 0.1, 0.4])
    rs = ReadSet()
    for k in range(nreads):
        a = rnd.randrange(0, npos - 1)
        span = pos[a:a + rnd.randrange(2, maxlen + 1)]
        if len(span) < 2:
            span = pos[-2:]
        r = Read("r%d" % k, 60, rnd.randrange(3), 0)
        for i, p in enumerate(span):
            if i in (0, len(span) - 1) or rnd.random() >= gap:
                r.add_variant(p, 0, [0.0, -5.0], rnd.choice(quals))
        rs.add(r)
    rs.sort()
    max_cov = rnd.choice([1, 2, 3, 5, 8, 15])
    pref = rnd.choice([None, None, [0], [1, 2], [7]])
    bridging = rnd.choice([True, True, False])
    sel = sorted(readselection(rs, max_cov, None if pref is None else set(pref), bridging))
    off, vpos, vq, src = [0], [], [], []
    for read in rs:
        for v in read:
            vpos.append(v.position)
            vq.append(v.quality)
        off.append(len(vpos))
        src.append(read.source_id)
    return dict(off=off, pos=vpos, qual=vq, src=src, max_cov=max_cov, pref=[-1] if pref is None else pref,
                bridging=int(bridging), sel=sel)


def main(n=160):
    cases = [make_case(1000 + s) for s in range(n)]
    out = {}
    for key, dt in (("off", np.uint64), ("pos", np.int32), ("qual", np.int32), ("src", np.int32), ("pref", np.int32),
                    ("sel", np.uint32)):
        out[key] = np.concatenate([np.asarray(c[key], dtype=dt) for c in cases])
        out[key + "_len"] = np.asarray([len(c[key]) for c in cases], dtype=np.int64)
    out["max_cov"] = np.asarray([c["max_cov"] for c in cases], dtype=np.int32)
    out["bridging"] = np.asarray([c["bridging"] for c in cases], dtype=np.int32)
    path = os.path.join(ROOT, "tests", "golden", "readselect", "cases.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", n, "cases,", int(out["off_len"].sum()) - n, "reads")


if __name__ == "__main__":
    main()
