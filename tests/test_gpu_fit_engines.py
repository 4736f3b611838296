"""The warp-cooperative fits -- exact (csrc/mdb_fit_warp.cuh, engine 2) and screened (csrc/mdb_fit_screen.cuh, engine 5) --
against the one-thread fit (csrc/mdb_compress.cuh), model by model: fit_next_model at EVERY start index of each series,
with and without a budget."""
import ctypes as C

import numpy as np
import pytest

from modelardb_rs_b200 import _native
from modelardb_rs_b200 import compression as mc
from tests.parity_cases import long_model_cases, small_cases

pytestmark = pytest.mark.gpu

DT = np.dtype([("start", np.uint32), ("end", np.uint32), ("min", np.uint32), ("max", np.uint32), ("last", np.uint32),
               ("bpv", np.uint32), ("type", np.int32), ("vlen", np.int32), ("aborted", np.int32), ("irregular", np.int32)])


def fit_models(ctx, ts, vals, eb, engine, starts, budget_ends):
    ts = np.ascontiguousarray(ts, np.int64)
    vals = np.ascontiguousarray(vals, np.float32)
    starts = np.ascontiguousarray(starts, np.uint32)
    budget_ends = np.ascontiguousarray(budget_ends, np.uint32)
    out = np.zeros(len(starts), DT)
    _native.check(_native.lib().mdbcu_debug_fit_models(ctx._h, ts.ctypes.data, vals.ctypes.data, len(ts), eb[0], eb[1], engine,
                                                       starts.ctypes.data, budget_ends.ctypes.data, len(starts), out.ctypes.data))
    return out


CASES = [c for c in small_cases() if len(c[1]) <= 16_000 and len(c[3]) == 2]


ENGINES = pytest.mark.parametrize("engine", [2, 5], ids=["exact", "screened"])


@ENGINES
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_warp_fit_equals_thread_fit_at_every_start(case, engine):
    name, ts, vals, off, ebs = case
    ctx = mc.default_context()
    n = len(ts)
    starts = np.arange(n, dtype=np.uint32)
    for budget in (None, 5, 40, 100):
        be = np.full(n, n, np.uint32) if budget is None else np.minimum(starts + budget, n).astype(np.uint32)
        a = fit_models(ctx, ts, vals, ebs[0], 1, starts, be)
        b = fit_models(ctx, ts, vals, ebs[0], engine, starts, be)
        assert np.array_equal(a["aborted"], b["aborted"]), (name, budget, np.flatnonzero(a["aborted"] != b["aborted"])[:5])
        ok = a["aborted"] == 0
        for f in ("start", "end", "min", "max", "last", "bpv", "type", "vlen"):
            bad = np.flatnonzero(ok & (a[f] != b[f]))
            assert len(bad) == 0, f"{name} budget={budget}: field {f} differs at starts {bad[:5]}: {a[bad[:3]]} vs {b[bad[:3]]}"


# ---- long models: the wide steps of the warp engine (mdb_fit_warp.cuh) --------------------------------------------

LONG_CASES = long_model_cases()


@ENGINES
@pytest.mark.parametrize("case", LONG_CASES, ids=[c[0] for c in LONG_CASES])
def test_warp_fit_equals_thread_fit_on_long_models(case, engine):
    name, ts, vals, eb = case
    ctx = mc.default_context()
    n = len(ts)
    rng = np.random.default_rng(7)
    starts = np.unique(np.concatenate([np.arange(0, 40), rng.integers(0, n, 40), np.arange(6_990, 7_003), np.arange(8_995, 9_005),
                                       np.arange(12_340, 12_350), np.arange(n - 600, n, 37)])).astype(np.uint32)
    for budget in (None, 700, 3_000):
        be = np.full(len(starts), n, np.uint32) if budget is None else np.minimum(starts + budget, n).astype(np.uint32)
        a = fit_models(ctx, ts, vals, eb, 1, starts, be)
        b = fit_models(ctx, ts, vals, eb, engine, starts, be)
        assert np.array_equal(a["aborted"], b["aborted"]), (name, budget, starts[np.flatnonzero(a["aborted"] != b["aborted"])[:5]])
        ok = a["aborted"] == 0
        for f in ("start", "end", "min", "max", "last", "bpv", "type", "vlen", "irregular"):
            bad = np.flatnonzero(ok & (a[f] != b[f]))
            assert len(bad) == 0, f"{name} budget={budget}: field {f} differs at starts {starts[bad[:5]]}: {a[bad[:3]]} vs {b[bad[:3]]}"


@pytest.mark.parametrize("fit_engine", [3, 5], ids=["exact", "screened"])
@pytest.mark.parametrize("chunk_len", [0, 512, 4096])
def test_long_model_series_match_oracle(oracle, chunk_len, fit_engine):
    """The same series through the whole compress path (chunk speculation with budgets) against the oracle."""
    from tests.parity_cases import assert_segments_equal
    ctx = mc.Context(0)
    ctx.set_chunk_len(chunk_len)
    ctx.set_fit_engine(fit_engine)
    by_eb = {}
    for name, ts, vals, eb in LONG_CASES:
        by_eb.setdefault(eb, []).append((name, ts, vals))
    for eb, group in by_eb.items():
        ts = np.concatenate([g[1] for g in group])
        vals = np.concatenate([g[2] for g in group])
        off = np.arange(len(group) + 1, dtype=np.uint64) * np.uint64(len(group[0][1]))
        want = oracle.compress(ts, vals, off, eb=eb, n_threads=8)
        got = mc.compress(ts, vals, off, mc.ErrorBound(*eb), ctx).to_host()
        assert_segments_equal(got, want, f"long models eb={eb} chunk_len={chunk_len}: {[g[0] for g in group]}")
    ctx.close()


# ---- the data the screened engine is meant for: noisy series, timestamps at epoch scale -----------------------------

SINES = [  # base, amplitude, period, noise, first timestamp, interval, bound (the cases of tests/test_screen_fit_emulated.py)
    (100.0, 10.0, 1000.0, 0.1, 1_700_000_000_000, 1, (2, 1.0)),
    (60.0, 19.0, 520.0, 0.1, 1_700_000_000_000, 1, (2, 1.0)),
    (140.0, 1.5, 1900.0, 0.1, 1_700_000_000_000, 1, (2, 1.0)),
    (100.0, 10.0, 1000.0, 0.1, 0, 1, (2, 1.0)),
    (100.0, 10.0, 1000.0, 0.1, 1_700_000_000_000_000, 1000, (2, 5.0)),
    (100.0, 10.0, 1000.0, 0.1, -9_000_000_000_000_000, 3_600_000, (1, 0.5)),
    (0.0, 10.0, 300.0, 0.05, 1_700_000_000_000, 100, (2, 10.0)),
    (1e-30, 1e-31, 700.0, 1e-33, 1_700_000_000_000, 1, (2, 1.0)),
    (100.0, 10.0, 1000.0, 0.0, 1_700_000_000_000, 1, (2, 1.0)),
]


@pytest.mark.parametrize("k", range(len(SINES)))
def test_screened_fit_equals_thread_fit_on_sine_series(k):
    """fit_next_model at 4000 starts of a sine + noise series: the screen decides most comparisons in f32 (with the device's
    own reciprocal), the doubtful ones in f64; every field must equal the one-thread fit's."""
    base, amp, period, noise, t0, step, eb = SINES[k]
    n = 40_000
    rng = np.random.default_rng(7 + k)
    i = np.arange(n)
    vals = (base + amp * np.sin(2 * np.pi * i / period + 1.0) + noise * rng.standard_normal(n)).astype(np.float32)
    ts = (t0 + step * i).astype(np.int64)
    ctx = mc.default_context()
    starts = np.unique(np.concatenate([np.arange(0, 2000), rng.integers(0, n, 2000)])).astype(np.uint32)
    for budget in (None, 600):
        be = np.full(len(starts), n, np.uint32) if budget is None else np.minimum(starts + budget, n).astype(np.uint32)
        a = fit_models(ctx, ts, vals, eb, 1, starts, be)
        b = fit_models(ctx, ts, vals, eb, 5, starts, be)
        assert np.array_equal(a["aborted"], b["aborted"]), (k, budget)
        ok = a["aborted"] == 0
        for f in ("start", "end", "min", "max", "last", "bpv", "type", "vlen", "irregular"):
            bad = np.flatnonzero(ok & (a[f] != b[f]))
            assert len(bad) == 0, f"sine {k} budget={budget}: field {f} differs at starts {starts[bad[:5]]}: {a[bad[:3]]} vs {b[bad[:3]]}"


@pytest.mark.parametrize("fit_engine", [3, 5], ids=["exact", "screened"])
def test_sine_series_match_oracle(oracle, fit_engine):
    """The same series as units of one batch through the whole compress path, per bound, against the oracle."""
    from tests.parity_cases import assert_segments_equal
    n = 150_000
    ctx = mc.Context(0)
    ctx.set_fit_engine(fit_engine)
    for k, (base, amp, period, noise, t0, step, eb) in enumerate(SINES):
        rng = np.random.default_rng(70 + k)
        i = np.arange(n)
        vals = np.concatenate([(base + amp * np.sin(2 * np.pi * i / period + ph) + noise * rng.standard_normal(n)).astype(np.float32)
                               for ph in (0.0, 2.0)])
        ts = np.tile((t0 + step * i).astype(np.int64), 2)
        off = np.array([0, n, 2 * n], np.uint64)
        want = oracle.compress(ts, vals, off, eb=eb, n_threads=8)
        got = mc.compress(ts, vals, off, mc.ErrorBound(*eb), ctx).to_host()
        assert_segments_equal(got, want, f"sine {k} fit_engine={fit_engine}")
    ctx.close()
