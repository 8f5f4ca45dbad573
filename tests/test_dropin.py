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
