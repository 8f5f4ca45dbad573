"""Checkpoint schedule (giggles_b200/csrc/schedule.cpp) — host only.  gg_schedule_dry_run builds the
launch list for a given HBM budget and replays it symbolically (every op must find the column it
expects where it expects it)."""
import numpy as np
import pytest

from giggles_b200 import capi


def test_everything_fits_single_level():
    cb = [1 << 20] * 64
    fb = [4096] * 64
    r = capi.schedule_dry_run(cb, fb, 1 << 30)
    assert r["levels"] == 1 and r["n_fwd"] == 64 and r["n_bwd"] == 64     # 63 steps + the all-ones last column


@pytest.mark.parametrize("ncols,budget_cols", [(100, 40), (100, 25), (1000, 80), (1000, 40), (5000, 60), (257, 33)])
def test_checkpointed_schedules_are_valid_and_bounded(ncols, budget_cols):
    col = 1 << 18
    r = capi.schedule_dry_run([col] * ncols, [1024] * ncols, budget_cols * col)
    assert r["n_fwd"] == ncols
    assert r["arena_bytes"] <= budget_cols * col
    assert r["levels"] >= 2
    # every extra level costs at most one more backward pass
    assert r["n_bwd"] <= r["levels"] * ncols


def test_variable_column_sizes():
    rng = np.random.default_rng(0)
    u = rng.integers(0, 9, size=400)
    cb = ((1 << u) * 1000 * 4).astype(np.uint64)
    fb = ((1 << u) * 6 * 4).astype(np.uint64)
    total = int(cb.sum())
    for frac in (2.0, 0.5, 0.2, 0.1):
        r = capi.schedule_dry_run(cb, fb, int(total * frac) + 4 * int(cb.max()))
        assert r["n_fwd"] == 400


def test_budget_too_small_is_an_error():
    with pytest.raises(capi.GGError) as ei:
        capi.schedule_dry_run([1 << 20] * 100, [1024] * 100, 3 << 20)
    assert ei.value.status == capi.GG_ERR_NOMEM


def test_single_and_empty():
    assert capi.schedule_dry_run([1000], [100], 1 << 20)["n_fwd"] == 1
    assert capi.schedule_dry_run([], [], 1 << 20)["n_fwd"] == 0


def test_columns_of_a_fifth_of_hbm_use_the_quadratic_fallback():
    # configs[3] with R = 88: one column is 33 GB; only two rolling backward columns fit beside alpha and the last beta
    col, fg = 33_235_664_896, 50_331_648
    r = capi.schedule_dry_run([col] * 8, [fg] * 8, 171_000_000_000)
    assert r["n_fwd"] == 8 and r["n_bwd"] == 1 + 8 * 7 // 2 and r["arena_bytes"] <= 171_000_000_000
    assert r["levels"] >= 2            # a recomputing schedule never reports "everything fitted"
    with pytest.raises(capi.GGError):
        capi.schedule_dry_run([col] * 8, [fg] * 8, 140_000_000_000)
