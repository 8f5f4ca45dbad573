"""Runs the reference's own Python surface (giggles.core, built by integration/build_dropin.sh with the B200
GenotypeHMM shim) on golden fixtures, the way giggles/cli/genotype.py:288-330 drives it.  Executed in a
subprocess by tests/test_dropin.py with PYTHONPATH=integration/_build.  Prints one JSON line per fixture."""
import json
import sys

import numpy as np

from giggles.core import (Genotype, GenotypeHMM, NumericSampleIds, Pedigree, PhredGenotypeLikelihoods, Read,
                          ReadSet)


def run(path):
    d = np.load(path)
    pos, nall, aref, recomb = d["positions"], d["n_alleles"], d["allele_ref"], d["recomb"]
    off, ecol, eall = d["read_entry_offset"], d["entry_column"], d["entry_allele"]
    soff, sc, tag, order = d["entry_score_offset"], d["entry_scores"], d["read_tag"], d["read_order"]
    ids = NumericSampleIds()
    reads = []
    for k in range(len(tag)):
        r = Read(f"read{int(order[k]):05d}", 60, 0, ids["sample"], -1, "", int(d["reg_const"]), float(d["base_const"]))
        for e in range(off[k], off[k + 1]):
            r.add_variant(int(pos[ecol[e]]), int(eall[e]), [float(x) for x in sc[soff[e]:soff[e + 1]]], 30)
        r.add_haplotag({-1: "none", 0: "H1", 1: "H2"}[int(tag[k])], -1)
        reads.append(r)
    rs = ReadSet()
    for k in np.random.default_rng(1).permutation(len(reads)):
        rs.add(reads[int(k)])
    rs.sort()
    C = len(pos)
    ped = Pedigree(ids)
    ped.add_individual("sample", [Genotype([]) for _ in range(C)],
                       [PhredGenotypeLikelihoods([1.0 / 3] * 3, 2, 2) for _ in range(C)])
    hmm = GenotypeHMM(ids, rs, [float(x) for x in recomb], ped, aref.shape[1], [int(x) for x in pos],
                      [int(x) for x in nall], [[int(a) for a in row] for row in aref])
    like = []
    for c in range(C):
        gl = hmm.get_genotype_likelihoods("sample", c, int(nall[c]))
        assert isinstance(gl, PhredGenotypeLikelihoods)
        like.extend(list(gl))
    like = np.array(like)
    want = d["likelihoods"]
    big = want >= 1e-3
    rel = float(np.max(np.abs(like[big] - want[big]) / want[big])) if big.any() else 0.0
    ab = float(np.max(np.abs(like[~big] - want[~big]))) if (~big).any() else 0.0
    return {"fixture": path.split("/")[-1], "columns": int(C), "max_rel_big": rel, "max_abs_small": ab,
            "nan": int(np.isnan(like).sum())}


if __name__ == "__main__":
    for p in sys.argv[1:]:
        try:
            print(json.dumps(run(p)))
        except RuntimeError as e:
            print(json.dumps({"fixture": p.split("/")[-1], "error": str(e)}))
