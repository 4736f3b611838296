"""The warp-cooperative fit (csrc/mdb_fit_warp.cuh) against the one-thread fit (csrc/mdb_compress.cuh),
model by model: fit_next_model at EVERY start index of each series, with and without a budget."""
import ctypes as C

import numpy as np
import pytest

from modelardb_rs_b200 import _native
from modelardb_rs_b200 import compression as mc
from tests.parity_cases import small_cases

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
