"""Parity AT THE BENCH SHAPE: the pipelined step kernel with many work items per CTA, checked against the CPU
oracle (not against another mode of the same engine).

The grid of `step_tma_kernel` is min(2^u, 296) persistent CTAs, so only columns with u >= 10 make a CTA walk several
items: ring wrap-around and parity flips across items, the one-item-ahead F/G prefetch, the L2 look-ahead of the
forward step (needs `far < n_items`), the double-buffered column-sum scratch.  The oracle is O(states) per column
(about 12 M states/s on one core), so these problems are a handful of columns each.  Reference functions under test:
src/genotypehmm.cpp:249-323 (backward), :412-509 (forward), src/column.cpp:72-122 (compatible bipartitions).
"""
import numpy as np
import pytest

import oracle
from conftest import assert_posteriors_close
from giggles_b200 import capi
from giggles_b200.synth import make_problem

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _need_device(have_gpu):
    if not have_gpu:
        pytest.fail("no CUDA device: the GenotypeHMM path has no CPU fallback")


# Weak evidence per read (emission ratio 1.1 per unit of edit distance instead of e, 80 % of the entries support the
# true allele): with ten or more reads per column the default scores saturate every posterior at 0 or 1 and a wrong
# block among thousands would go unnoticed; these leave most columns with posteriors well inside (0, 1).
SOFT = dict(base_const=1.1, accuracy=0.8)

# (name, make_problem kwargs, minimum max-u the problem must reach, kernel variants)
SHAPES = [
    # short reads: several reads start and end inside the window (reduce sets of 2 and 4 blocks, broadcast bits)
    ("R88_u11_events", dict(n_columns=8, n_haps=88, region_bp=8 * 124, coverage=40, max_coverage=11, read_len_mean=500,
                            read_len_sigma=0.5, window=True, seed=1101, **SOFT), 10, (0, 2, 3)),
    # long reads: block-local steps, 14 items per CTA, look-ahead live in every launch
    ("R88_u12_window", dict(n_columns=5, n_haps=88, region_bp=5 * 124, coverage=30, max_coverage=12, window=True,
                            seed=1102, **SOFT), 12, (0,)),
    ("R88_u10_mixed_tags_multiallelic", dict(n_columns=8, n_haps=88, region_bp=8 * 124, coverage=60, max_coverage=20,
                                             read_len_mean=600, read_len_sigma=0.5, tags="mixed", window=True,
                                             multiallelic_frac=0.4, max_alleles=6, missing_frac=0.02, seed=1103, **SOFT), 9, (0, 3)),
    ("R94_u10_events", dict(n_columns=8, n_haps=94, region_bp=8 * 124, coverage=40, max_coverage=10, read_len_mean=500,
                            read_len_sigma=0.5, window=True, seed=1104, **SOFT), 9, (0, 2)),       # float2 rows (R % 4 == 2)
    ("R128_u10_events", dict(n_columns=6, n_haps=128, region_bp=6 * 124, coverage=40, max_coverage=10, read_len_mean=400,
                             read_len_sigma=0.5, window=True, seed=1105, **SOFT), 9, (0, 2)),
    # wide panels (hmm_wide_kernel.cuh: warps own columns; 8 = the row-tiled pipelined kernel instead): configs[4]'s
    # 400 haplotypes with 3..16 alleles per site, more blocks than SMs, reads starting and ending inside the window
    ("R400_u8_multiallelic_events", dict(n_columns=5, n_haps=400, region_bp=5 * 124, coverage=40, max_coverage=8,
                                         read_len_mean=350, read_len_sigma=0.5, window=True, multiallelic_frac=1.0,
                                         max_alleles=16, seed=1106, **SOFT), 8, (0, 8)),
    ("R200_u9_multiallelic_mixed_tags", dict(n_columns=6, n_haps=200, region_bp=6 * 124, coverage=50, max_coverage=12,
                                             read_len_mean=400, read_len_sigma=0.5, tags="mixed", window=True,
                                             multiallelic_frac=0.5, max_alleles=8, missing_frac=0.02, seed=1107, **SOFT), 7, (0, 8)),
    ("R132_u9_events", dict(n_columns=6, n_haps=132, region_bp=6 * 124, coverage=40, max_coverage=9, read_len_mean=400,
                            read_len_sigma=0.5, window=True, seed=1108, **SOFT), 8, (0,)),          # 33 float4 per row: 3 column groups, the last one a single lane pair
    ("R256_u8_events", dict(n_columns=5, n_haps=256, region_bp=5 * 124, coverage=40, max_coverage=8, read_len_mean=400,
                            read_len_sigma=0.5, window=True, seed=1109, **SOFT), 7, (0,)),
]


@pytest.mark.parametrize("name,kw,min_u,variants", SHAPES, ids=[s[0] for s in SHAPES])
def test_many_items_per_cta_against_oracle(name, kw, min_u, variants):
    p = make_problem(**kw)
    u = p.untagged_per_column()
    assert int(u.max()) >= min_u, u.tolist()          # 2^u >> 296 CTAs
    want = oracle.oracle_run(p)
    assert np.isfinite(want).all()
    assert np.mean((want > 1e-4) & (want < 0.999)) > 0.3          # not saturated: the comparison discriminates
    for kv in variants:
        got = capi.hmm_run(p, device=0, kernel_variant=kv).values
        assert_posteriors_close(got, want)
    # a tight budget forces recompute launches at this size too; results must not move by a bit
    base = capi.hmm_run(p, device=0).values
    h = capi.Hmm(p, device=0)
    full = h.stats()["state_bytes_reserved"]
    h.close()
    for frac in (0.5, 0.7):
        try:
            h = capi.Hmm(p, device=0, budget=int(full * frac))
        except capi.GGError as e:
            assert e.status == capi.GG_ERR_NOMEM
            continue
        h.sweep()
        assert np.array_equal(h.fetch().values, base)
        h.close()
        break


def test_u_changes_inside_the_window_really_happen():
    # guard for the cases above: the "events" shapes must contain reads that start and reads that end inside
    p = make_problem(**SHAPES[0][1])
    u = p.untagged_per_column()
    assert len(set(u.tolist())) >= 2
    pv = capi.PlanView(p)
    started = ended = 0
    for c in range(1, p.n_columns):
        a, b = set(pv.untagged(c - 1)), set(pv.untagged(c))
        started += len(b - a)
        ended += len(a - b)
    assert started >= 2 and ended >= 2


_COLOWN_CASE = r"""
import sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/tests")
import numpy as np
from giggles_b200 import capi
from giggles_b200.synth import make_problem
import test_gpu_bench_shape as T
out = {}
for name, kw, min_u, variants in T.SHAPES:
    if name in T.COLOWN_SHAPES:
        out[name] = capi.hmm_run(make_problem(**kw), device=0).values
np.savez(sys.argv[2], **out)
print("ok")
"""

COLOWN_SHAPES = ("R88_u11_events",)     # R = 128 runs the default mode in the matrix above


def test_column_owning_kernel_modes_against_oracle(tmp_path):
    """GG_COLOWN picks the step kernel of mid-size panels per direction (0: step_tma_kernel for both, 1: column-owning
    backward step — the default —, 3: column-owning for both).  Every mode against the oracle at the bench shape (the
    knob is read once per process, hence subprocesses; the oracle runs once, here)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    want = {name: oracle.oracle_run(make_problem(**kw)) for name, kw, _, _ in SHAPES if name in COLOWN_SHAPES}
    for mode in ("0", "1", "3"):
        path = str(tmp_path / f"colown{mode}.npz")
        out = subprocess.run([sys.executable, "-c", _COLOWN_CASE, root, path], env=dict(os.environ, GG_COLOWN=mode),
                             capture_output=True, text=True, timeout=900)
        assert out.returncode == 0, (mode, out.stderr[-2000:])
        got = np.load(path)
        for name in COLOWN_SHAPES:
            assert_posteriors_close(got[name], want[name])
