"""Generates tests/golden/editdist_golden.json from the reference's own Cython module giggles.align
(oracle/_ref/pyref, built by `bash oracle/build_ref.sh --python`): seeded pairs in the shapes the realignment
produces (a query against ref / alt alleles between identical overhangs, giggles/variants.py:207-239) plus the edge
cases of edit_distance(s, t, maxdiff) (giggles/align.pyx:15-107).  Run here, in the container that has /root/reference."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402


def cases(seed=7):
    rng = np.random.default_rng(seed)

    def dna(n, alphabet="ACGT"):
        return "".join(rng.choice(list(alphabet), size=n)) if n else ""

    def mutate(s, rate):
        out = []
        for ch in s:
            r = rng.random()
            if r < rate / 3:
                continue
            if r < 2 * rate / 3:
                out.append(str(rng.choice(list("ACGT"))))
                continue
            out.append(ch)
            if r < rate:
                out.append(str(rng.choice(list("ACGT"))))
        return "".join(out)

    pairs = [("", ""), ("", "ACGT"), ("ACGT", ""), ("A", "A"), ("A", "C"), ("kitten", "sitting"), ("ACGT" * 20, "ACGT" * 20),
             ("AAAA", "AAAAAAAA"), ("ACGTN*", "NNAC*"), ("äöü", "aou"), ("A" * 64, "C" * 64), ("A" * 65, "A" * 63 + "CG")]
    # realignment shapes: overhang + allele + overhang against a noisy query
    for _ in range(120):
        over = int(rng.choice([10, 30, 100]))
        left, right = dna(over), dna(over)
        ref = dna(int(rng.choice([0, 1, 5, 50, 300, 1200])))
        alt = dna(int(rng.choice([0, 1, 8, 60, 500, 2500])))
        truth = ref if rng.random() < 0.5 else alt
        query = mutate(left + truth + right, float(rng.choice([0.0, 0.01, 0.05, 0.15])))
        pairs.append((query, left + ref + right))
        pairs.append((query, left + alt + right))
    # long patterns: more than 2048 characters after stripping (several word groups), and 64-character boundaries
    for n in (63, 64, 65, 127, 128, 129, 2047, 2048, 2049, 4100, 6500):
        a = dna(n)
        pairs.append((a, mutate(a, 0.1)))
        pairs.append((dna(n), dna(max(1, n - 17))))
    # a wide alphabet (more than 15 distinct bytes) and lower case
    pairs.append(("The quick brown fox jumps over the lazy dog", "the quick brown dog jumps over the lazy fox!"))
    pairs.append((dna(300, "acgtn"), dna(280, "acgtn")))
    out = []
    ed = oracle.reference_edit_distance()
    assert ed is not None, "oracle/_ref/pyref/giggles/align missing: run `bash oracle/build_ref.sh --python`"
    for s, t in pairs:
        d = {}
        for maxdiff in (-1, 0, 1, 3, 10, 50):
            if maxdiff != -1 and len(s) * len(t) > 4_000_000:
                continue
            d[str(maxdiff)] = int(ed(s, t, maxdiff))
        out.append([s, t, d])        # one entry per pair: {maxdiff: distance}
    return out


if __name__ == "__main__":
    c = cases()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "editdist_golden.json")
    with open(path, "w") as f:
        json.dump(c, f)
    print(f"{sum(len(x[2]) for x in c)} cases on {len(c)} pairs -> {path} ({os.path.getsize(path) / 1e6:.2f} MB)")
