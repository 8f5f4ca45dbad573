import glob
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # the product library and the CPU oracle are built in-tree; build them if a fresh checkout lacks them
    lib = os.path.join(ROOT, "giggles_b200", "libgiggles_b200.so")
    if not os.path.exists(lib):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "giggles_b200", "csrc")])
    ora = os.path.join(ROOT, "oracle", "_build", "libhmm_oracle.so")
    if not os.path.exists(ora):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "_build/libhmm_oracle.so"])


def golden_files():
    return sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz")))


def load_golden(path):
    from giggles_b200.problem import Problem
    d = np.load(path)
    return Problem.from_npz_dict(d), d["likelihoods"]


@pytest.fixture(scope="session")
def have_gpu():
    from giggles_b200 import capi
    return capi.lib().gg_device_count() > 0


# ---- the parity bar: posteriors rel 1e-5 for p >= 1e-3, abs 2e-6 below; same GT; GQ +-1
def assert_posteriors_close(got, want, rel=1e-5, abs_small=2e-6):
    got = np.asarray(got)
    want = np.asarray(want)
    assert got.shape == want.shape
    assert not np.isnan(got).any()
    big = want >= 1e-3
    if big.any():
        r = np.abs(got[big] - want[big]) / want[big]
        assert r.max() <= rel, f"max rel err {r.max():.3e} on posteriors >= 1e-3"
    if (~big).any():
        a = np.abs(got[~big] - want[~big])
        assert a.max() <= abs_small, f"max abs err {a.max():.3e} on posteriors < 1e-3"


def py_determine_genotype(l, threshold=0.0):
    """Mirror of reference giggles/cli/genotype.py:82-95 on a plain list (index of the call or -1)."""
    order = sorted(range(len(l)), key=lambda i: l[i])
    if len(l) >= 2 and l[order[-1]] > l[order[-2]] and l[order[-1]] - l[order[-2]] > threshold:
        return order[-1]
    return -1


def py_gq(l, idx):
    """reference giggles/vcf.py:850-872."""
    import math
    if idx < 0:
        return -1
    q = sum(x for i, x in enumerate(l) if i != idx)
    return min(round(-10.0 * math.log10(q)), 10000) if q > 0 else 10000
