"""The drop-in boundary end to end: the REFERENCE's Cython module giggles.core, rebuilt with
integration/genotypehmm.{h,cpp} in place of src/genotypehmm.{h,cpp} and linked against libgiggles_b200.so,
driven through the reference's own Python API.  The module is built in the build container
(integration/build_dropin.sh needs the reference checkout) and travels to the GPU box as a built artefact."""
import glob
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "integration", "_build")
HAVE = bool(glob.glob(os.path.join(BUILD, "giggles", "core*.so")))
needs_module = pytest.mark.skipif(not HAVE, reason="integration/_build not built (needs the reference checkout)")


def _drive(fixtures):
    env = dict(os.environ, PYTHONPATH=BUILD)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "dropin_driver.py")] + fixtures, env=env,
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    return [json.loads(l) for l in out.stdout.splitlines() if l.startswith("{")]


@needs_module
def test_dropin_fails_loudly_without_a_device(have_gpu):
    if have_gpu:
        pytest.skip("a device is present")
    res = _drive([os.path.join(ROOT, "tests", "golden", "untagged_small.npz")])
    assert "no CUDA device" in res[0]["error"]          # RuntimeError through Cython's `except +`, no CPU fallback


@needs_module
@pytest.mark.gpu
def test_dropin_matches_reference_golden():
    fixtures = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz")))
    res = _drive(fixtures)
    assert len(res) == len(fixtures)
    for r in res:
        assert "error" not in r, r
        assert r["nan"] == 0
        assert r["max_rel_big"] <= 1e-5 and r["max_abs_small"] <= 2e-6, r


_READSELECT = r"""
import json, sys
import numpy as np
from giggles.core import Read, ReadSet, readselection
d = np.load(sys.argv[1])
cur = {k: 0 for k in ("off", "pos", "qual", "src", "pref", "sel")}
bad = []
for i in range(len(d["max_cov"])):
    c = {}
    for k in cur:
        n = int(d[k + "_len"][i]); c[k] = d[k][cur[k]:cur[k] + n].tolist(); cur[k] += n
    rs = ReadSet()
    for k in range(len(c["src"])):
        r = Read("r%d" % k, 60, c["src"][k], 0)
        for e in range(c["off"][k], c["off"][k + 1]):
            r.add_variant(c["pos"][e], 0, [0.0, -5.0], c["qual"][e])
        rs.add(r)
    pref = None if c["pref"] == [-1] else set(c["pref"])
    sel = readselection(rs, int(d["max_cov"][i]), pref, bool(d["bridging"][i]))
    if not isinstance(sel, set) or sorted(sel) != c["sel"]:
        bad.append(i)
print(json.dumps({"cases": int(len(d["max_cov"])), "bad": bad}))
"""


@needs_module
def test_dropin_readselection_matches_reference_golden():
    """giggles.core.readselection of the drop-in module (integration/readselect.pyx -> gg_readselection) through the
    reference's own Read / ReadSet classes; host only, so this one also runs without a GPU."""
    env = dict(os.environ, PYTHONPATH=BUILD)
    out = subprocess.run([sys.executable, "-c", _READSELECT, os.path.join(ROOT, "tests", "golden", "readselect", "cases.npz")],
                         env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    res = json.loads(out.stdout.splitlines()[-1])
    assert res == {"cases": 160, "bad": []}
