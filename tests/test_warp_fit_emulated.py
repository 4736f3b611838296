"""The warp-cooperative fit engine (modelardb_rs_b200/csrc/mdb_fit_warp.cuh) run on the HOST: the 32 lanes of a warp are
32 cooperative fibers and every warp primitive is a rendezvous (tests/emu/warp_emu.h).  A debugging harness for the
GPU-less build container -- the same comparisons run on the device in tests/test_gpu_fit_engines.py -- that lets changes to
the engine be checked here before GPU time is spent: model by model against the one-thread fit, and through the whole
chunked compress (rounds and asynchronous scheduler, skip_rejected included) against the oracle."""
import numpy as np
import pytest

from tests import emu_lib as emu
from tests.parity_cases import assert_segments_equal, long_model_cases, small_cases

CASES = [c for c in small_cases() if len(c[3]) == 2]  # single-unit cases
FIELDS = ("start", "end", "min", "max", "last", "bpv", "type", "vlen", "irregular")


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_emulated_warp_fit_equals_thread_fit(case):
    name, ts, vals, off, ebs = case
    n = len(ts)
    rng = np.random.default_rng(3)
    starts = np.unique(np.concatenate([np.arange(0, min(n, 48)), rng.integers(0, n, 60), np.arange(max(0, n - 20), n)])).astype(np.uint32)
    for budget in (None, 40, 300):
        be = np.full(len(starts), n, np.uint32) if budget is None else np.minimum(starts + budget, n).astype(np.uint32)
        a = emu.fit_models(ts, vals, ebs[0], 1, starts, be)
        b = emu.fit_models(ts, vals, ebs[0], 2, starts, be)
        assert np.array_equal(a["aborted"], b["aborted"]), (name, budget)
        ok = a["aborted"] == 0
        for f in FIELDS:
            bad = np.flatnonzero(ok & (a[f] != b[f]))
            assert len(bad) == 0, f"{name} budget={budget}: {f} differs at starts {starts[bad[:5]]}: {a[bad[:3]]} vs {b[bad[:3]]}"
    assert emu.division_mismatches() == 0  # the branch-free division sequence returned a / b every time


@pytest.mark.parametrize("chunk_len,sched", [(64, (0, 0)), (1000, (5, 3))], ids=["rounds-64", "async-1000"])
@pytest.mark.parametrize("case", [c for c in small_cases() if len(c[1]) <= 8000], ids=lambda c: c[0])
def test_emulated_warp_engine_compress_matches_oracle(oracle, case, chunk_len, sched):
    name, ts, vals, off, ebs = case
    want = oracle.compress(ts, vals, off, eb=ebs)
    got = emu.compress(ts, vals, off, eb=ebs, chunk_len=chunk_len, sched_seed=sched[0], in_flight=sched[1], engine=2)
    assert_segments_equal(got, want, f"{name} chunk_len={chunk_len} sched={sched}")


def test_emulated_wide_steps_equal_thread_fit():
    """The optional wide steps (512 points at once inside long models, compiled out of the product by default:
    MDB_FIT_WIDE_ENABLED) stay exact: a build with them enabled against the one-thread fit on the long-model series."""
    wide = emu.variant("MDB_FIT_WIDE_ENABLED=1")
    for name, ts, vals, eb in long_model_cases():
        ts, vals = np.ascontiguousarray(ts[:6000]), np.ascontiguousarray(vals[:6000])
        n = len(ts)
        starts = np.array([0, 1, 5, 100, 2999, 3000, 4000, 5000, 5990], np.uint32)
        for budget in (None, 700):
            be = np.full(len(starts), n, np.uint32) if budget is None else np.minimum(starts + budget, n).astype(np.uint32)
            a = emu.fit_models(ts, vals, eb, 1, starts, be, library=wide)
            b = emu.fit_models(ts, vals, eb, 2, starts, be, library=wide)
            assert np.array_equal(a["aborted"], b["aborted"]), (name, budget)
            ok = a["aborted"] == 0
            for f in FIELDS:
                assert np.array_equal(a[f][ok], b[f][ok]), (name, budget, f)
    assert wide.emu_division_mismatches() == 0


def _fuzz_series(rng):
    """A short series stitched from constants, ramps, noise, repeats of one value, signed zeros and special values."""
    parts = []
    for _ in range(int(rng.integers(2, 7))):
        kind = int(rng.integers(0, 7))
        m = int(rng.integers(1, 400))
        scale = float(10.0 ** rng.integers(-3, 6))
        if kind == 0:
            parts.append(np.full(m, rng.normal() * scale))
        elif kind == 1:
            parts.append(rng.normal() * scale + rng.normal() * scale * 1e-3 * np.arange(m))
        elif kind == 2:
            parts.append(rng.normal() * scale + rng.standard_normal(m) * scale * 10.0 ** rng.integers(-4, 0))
        elif kind == 3:
            parts.append(np.repeat(rng.standard_normal(max(1, m // 9)) * scale, 9)[:m])
        elif kind == 4:
            parts.append(np.where(rng.random(m) < 0.5, 0.0, -0.0))
        elif kind == 5:
            parts.append(rng.choice([np.nan, np.inf, -np.inf, 1.0, 3.4e38, -3.4e38, 1e-45], m))
        else:
            parts.append(np.cumsum(rng.standard_normal(m)) * scale)
    return np.concatenate(parts).astype(np.float32)


@pytest.mark.parametrize("seed", range(64))
def test_emulated_warp_engine_fuzz(oracle, seed):
    """Random stitched series, random bounds, regular or irregular timestamps: the warp engine's compress (asynchronous
    scheduler, random interleaving) is the oracle's, column for column."""
    rng = np.random.default_rng(1000 + seed)
    vals = _fuzz_series(rng)
    n = len(vals)
    step = rng.integers(1, 2000, n) if seed % 3 == 0 else np.full(n, int(rng.integers(1, 5000)))
    ts = (int(rng.integers(0, 2_000_000_000_000_000)) + np.cumsum(step)).astype(np.int64)
    eb = [(0, 0.0), (1, float(10.0 ** rng.integers(-3, 3))), (2, float(rng.choice([0.01, 0.5, 1.0, 5.0, 30.0, 100.0])))][seed % 3 if seed % 5 else 0]
    want = oracle.compress(ts, vals, eb=eb)
    got = emu.compress(ts, vals, eb=eb, chunk_len=int(rng.choice([8, 100, 700])), sched_seed=seed + 1, in_flight=int(rng.choice([1, 2, 9])), engine=2)
    assert_segments_equal(got, want, f"fuzz seed={seed} eb={eb} n={n}")


@pytest.mark.parametrize("chunk_len,sched", [(4096, (3, 4)), (700, (9, 2))], ids=["async-4096", "async-700"])
def test_emulated_warp_engine_on_long_models(oracle, chunk_len, sched):
    """The long-model series of the GPU tests (models of thousands of points, values right around the bounds, zeros, sign
    changes, special values, extreme magnitudes) through the emulated warp engine and the asynchronous scheduler."""
    for name, ts, vals, eb in long_model_cases():
        want = oracle.compress(ts, vals, eb=eb)
        got = emu.compress(ts, vals, eb=eb, chunk_len=chunk_len, sched_seed=sched[0], in_flight=sched[1], engine=2)
        assert_segments_equal(got, want, f"{name} chunk_len={chunk_len}")
    assert emu.division_mismatches() == 0
