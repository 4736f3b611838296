"""The warp-cooperative fit (csrc/mdb_fit_warp.cuh) against the one-thread fit (csrc/mdb_compress.cuh),
model by model: fit_next_model at EVERY start index of each series, with and without a budget."""
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


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_warp_fit_equals_thread_fit_at_every_start(case):
    name, ts, vals, off, ebs = case
    ctx = mc.default_context()
    n = len(ts)
    starts = np.arange(n, dtype=np.uint32)
    for budget in (None, 5, 40, 100):
        be = np.full(n, n, np.uint32) if budget is None else np.minimum(starts + budget, n).astype(np.uint32)
        a = fit_models(ctx, ts, vals, ebs[0], 1, starts, be)
        b = fit_models(ctx, ts, vals, ebs[0], 2, starts, be)
        assert np.array_equal(a["aborted"], b["aborted"]), (name, budget, np.flatnonzero(a["aborted"] != b["aborted"])[:5])
        ok = a["aborted"] == 0
        for f in ("start", "end", "min", "max", "last", "bpv", "type", "vlen"):
            bad = np.flatnonzero(ok & (a[f] != b[f]))
            assert len(bad) == 0, f"{name} budget={budget}: field {f} differs at starts {bad[:5]}: {a[bad[:3]]} vs {b[bad[:3]]}"


# ---- long models: the wide steps of the warp engine (mdb_fit_warp.cuh) --------------------------------------------

LONG_CASES = long_model_cases()


@pytest.mark.parametrize("case", LONG_CASES, ids=[c[0] for c in LONG_CASES])
def test_warp_fit_equals_thread_fit_on_long_models(case):
    name, ts, vals, eb = case
    ctx = mc.default_context()
    n = len(ts)
    rng = np.random.default_rng(7)
    starts = np.unique(np.concatenate([np.arange(0, 40), rng.integers(0, n, 40), np.arange(6_990, 7_003), np.arange(8_995, 9_005),
                                       np.arange(12_340, 12_350), np.arange(n - 600, n, 37)])).astype(np.uint32)
    for budget in (None, 700, 3_000):
        be = np.full(len(starts), n, np.uint32) if budget is None else np.minimum(starts + budget, n).astype(np.uint32)
        a = fit_models(ctx, ts, vals, eb, 1, starts, be)
        b = fit_models(ctx, ts, vals, eb, 2, starts, be)
        assert np.array_equal(a["aborted"], b["aborted"]), (name, budget, starts[np.flatnonzero(a["aborted"] != b["aborted"])[:5]])
        ok = a["aborted"] == 0
        for f in ("start", "end", "min", "max", "last", "bpv", "type", "vlen", "irregular"):
            bad = np.flatnonzero(ok & (a[f] != b[f]))
            assert len(bad) == 0, f"{name} budget={budget}: field {f} differs at starts {starts[bad[:5]]}: {a[bad[:3]]} vs {b[bad[:3]]}"


@pytest.mark.parametrize("chunk_len", [0, 512, 4096])
def test_long_model_series_match_oracle(oracle, chunk_len):
    """The same series through the whole compress path (chunk speculation with budgets) against the oracle."""
    from tests.parity_cases import assert_segments_equal
    ctx = mc.Context(0)
    ctx.set_chunk_len(chunk_len)
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
