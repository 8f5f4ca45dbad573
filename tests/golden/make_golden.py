"""Generates tests/golden/*.npz by running the UNMODIFIED reference through its own Python API.

Run in the build container (needs /root/reference):
    bash oracle/build_ref.sh --python && python tests/golden/make_golden.py

For every seeded case the script builds ``Read`` objects exactly like ``giggles genotype`` does
(reference giggles/cli/genotype.py:288-330, giggles/core.pyx:208-214), lets the reference's
``ReadSet.sort()`` order them (src/readset.cpp:44-53), runs the reference ``GenotypeHMM``
(giggles/core.pyx:474-506) and stores (a) the flat problem in the order the reference's sorted
ReadSet has it and (b) the posteriors ``get_genotype_likelihoods`` returned.  The committed .npz
files are the parity pin on the GPU box, where neither /root/reference nor this script's imports
exist.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref", "pyref"))

from giggles.core import (Genotype, GenotypeHMM, NumericSampleIds, Pedigree,  # noqa: E402  (the reference)
                          PhredGenotypeLikelihoods, Read, ReadSet)

from giggles_b200.problem import Problem  # noqa: E402
from giggles_b200.synth import make_problem  # noqa: E402

CASES = {
    # name: make_problem kwargs
    "untagged_small": dict(n_columns=30, n_haps=6, coverage=5, max_coverage=6, read_len_mean=2000, tags="none", seed=11),
    "mixed_tags": dict(n_columns=40, n_haps=5, coverage=6, max_coverage=6, read_len_mean=2000, tags="mixed", seed=12),
    "all_tagged": dict(n_columns=50, n_haps=8, coverage=6, max_coverage=6, read_len_mean=2000, tags="all", seed=13),
    "multiallelic": dict(n_columns=30, n_haps=7, coverage=5, max_coverage=5, read_len_mean=2000, multiallelic_frac=0.6,
                         max_alleles=6, seed=14),
    "missing_panel": dict(n_columns=30, n_haps=6, coverage=5, max_coverage=5, read_len_mean=2000, missing_frac=0.15, seed=15),
    "gaps_unknown": dict(n_columns=30, n_haps=6, coverage=5, max_coverage=5, read_len_mean=2500, gap_frac=0.3,
                         unknown_allele_frac=0.15, seed=16),
    "single_column": dict(n_columns=1, n_haps=5, coverage=4, max_coverage=4, read_len_mean=800, seed=17, window=True),
    "two_columns": dict(n_columns=2, n_haps=4, coverage=4, max_coverage=4, read_len_mean=800, seed=18, window=True),
    "three_columns_k1": dict(n_columns=3, n_haps=4, coverage=4, max_coverage=4, read_len_mean=800, seed=19),
    "no_reads": dict(n_columns=12, n_haps=6, coverage=0.0, seed=20),
    "sparse_cover": dict(n_columns=60, n_haps=5, coverage=0.8, max_coverage=3, read_len_mean=1200, seed=21),
    "odd_haps": dict(n_columns=25, n_haps=9, coverage=5, max_coverage=5, read_len_mean=1500, seed=22),
    "window_start": dict(n_columns=20, n_haps=6, coverage=8, max_coverage=7, read_len_mean=3000, window=True, seed=23),
    "config1_slice": dict(n_columns=120, n_haps=10, region_bp=24000, coverage=10.0, max_coverage=8, seed=24),
    "panel88_tagged": dict(n_columns=12, n_haps=88, region_bp=1500, coverage=6, max_coverage=6, read_len_mean=1200,
                           tags="all", seed=25),
    "panel88_untagged": dict(n_columns=8, n_haps=88, region_bp=1000, coverage=3, max_coverage=2, read_len_mean=1000,
                             seed=26),
    "long_chain_ckpt": dict(n_columns=300, n_haps=4, coverage=3, max_coverage=3, read_len_mean=1500, seed=27),
    "noisy_ont": dict(n_columns=40, n_haps=6, coverage=6, max_coverage=6, read_len_mean=4000, read_len_sigma=0.6,
                      accuracy=0.85, seed=28),
    "base2_reg6": dict(n_columns=30, n_haps=6, coverage=5, max_coverage=5, read_len_mean=2000, seed=29, reg_const=6,
                       base_const=2.0),
    "all16_alleles": dict(n_columns=10, n_haps=12, coverage=3, max_coverage=3, read_len_mean=1500, multiallelic_frac=1.0,
                          min_alleles_multi=16, max_alleles=16, seed=30),
}


def run_reference(p: Problem, shuffle_seed: int):
    """Feed ``p`` through the reference's Python API; returns the problem re-flattened in the order of the
    reference's sorted ReadSet and the posteriors."""
    rng = np.random.default_rng(shuffle_seed)
    ids = NumericSampleIds()
    sample = "sample"
    reads = []
    off = p.read_entry_offset
    for k in range(p.n_reads):
        r = Read(f"read{k:05d}", 60, 0, ids[sample], -1, "", p.reg_const, p.base_const)
        for e in range(off[k], off[k + 1]):
            sc = p.entry_scores[p.entry_score_offset[e]:p.entry_score_offset[e + 1]]
            r.add_variant(int(p.positions[p.entry_column[e]]), int(p.entry_allele[e]), [float(x) for x in sc], 30)
        r.add_haplotag({-1: "none", 0: "H1", 1: "H2"}[int(p.read_tag[k])], -1)
        reads.append(r)
    rs = ReadSet()
    for k in rng.permutation(p.n_reads):          # insertion order must not matter: ReadSet.sort() fixes it
        rs.add(reads[int(k)])
    rs.sort()
    C_ = p.n_columns
    ped = Pedigree(ids)
    ped.add_individual(sample, [Genotype([]) for _ in range(C_)],
                       [PhredGenotypeLikelihoods([1.0 / 3] * 3, 2, 2) for _ in range(C_)])
    recomb = [float(x) for x in p.recomb]
    hmm = GenotypeHMM(ids, rs, recomb, ped, p.n_haps, [int(x) for x in p.positions], [int(x) for x in p.n_alleles],
                      [[int(a) for a in row] for row in p.allele_ref])
    like = []
    for c in range(C_):
        like.extend(list(hmm.get_genotype_likelihoods(sample, c, int(p.n_alleles[c]))))
    # re-flatten in the reference's sorted order
    order = [int(r.name[4:]) for r in rs]
    col_of = {int(pos): c for c, pos in enumerate(p.positions)}
    tag, roff, ecol, eall, soff, sc = [], [0], [], [], [0], []
    for r, k in zip(rs, order):
        tag.append(int(p.read_tag[k]))
        for e in range(off[k], off[k + 1]):
            ecol.append(int(p.entry_column[e]))
            eall.append(int(p.entry_allele[e]))
            sc.extend(p.entry_scores[p.entry_score_offset[e]:p.entry_score_offset[e + 1]].tolist())
            soff.append(len(sc))
        roff.append(len(ecol))
        assert [col_of[v.position] for v in r] == ecol[roff[-2]:roff[-1]]
    q = Problem(positions=p.positions, n_alleles=p.n_alleles, allele_ref=p.allele_ref, recomb=p.recomb,
                read_tag=np.array(tag, np.int8), read_entry_offset=np.array(roff, np.uint32),
                entry_column=np.array(ecol, np.uint32), entry_allele=np.array(eall, np.int32),
                entry_score_offset=np.array(soff, np.uint64), entry_scores=np.array(sc, np.float64),
                reg_const=p.reg_const, base_const=p.base_const)
    return q, np.array(like, dtype=np.float64), np.array(order, dtype=np.int64)


def main():
    for name, kw in CASES.items():
        p = make_problem(**kw)
        q, like, order = run_reference(p, shuffle_seed=kw["seed"])
        assert like.shape[0] == int(q.gl_offsets()[-1])
        d = q.to_npz_dict()
        d["likelihoods"] = like
        d["read_order"] = order
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
        print(f"{name}: C={q.n_columns} R={q.n_haps} reads={q.n_reads} max_u={int(q.untagged_per_column().max()) if q.n_columns else 0} "
              f"nan={int(np.isnan(like).sum())} bytes={os.path.getsize(os.path.join(HERE, name + '.npz'))}")


if __name__ == "__main__":
    main()
