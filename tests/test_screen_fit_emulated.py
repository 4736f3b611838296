"""The screened fit engine (modelardb_rs_b200/csrc/mdb_fit_screen.cuh) run on the HOST through tests/emu/warp_emu.h, like
the exact cooperative engine in tests/test_warp_fit_emulated.py: model by model against the one-thread fit, and through the
whole chunked compress against the oracle.  A debugging harness for the GPU-less build container; the same comparisons
run on the device in tests/test_gpu_fit_engines.py and tests/test_gpu_parity.py."""
import numpy as np
import pytest

from tests import emu_lib as emu
from tests.parity_cases import assert_segments_equal, long_model_cases, small_cases
from tests.test_warp_fit_emulated import _fuzz_series

CASES = [c for c in small_cases() if len(c[3]) == 2]  # single-unit cases
FIELDS = ("start", "end", "min", "max", "last", "bpv", "type", "vlen", "irregular")
SCREEN = 5  # engine number of the screened fit in the emulator and in mdbcu_debug_fit_models


def _assert_same_fits(ts, vals, eb, starts, budgets, label):
    n = len(ts)
    for budget in budgets:
        be = np.full(len(starts), n, np.uint32) if budget is None else np.minimum(starts + budget, n).astype(np.uint32)
        a = emu.fit_models(ts, vals, eb, 1, starts, be)
        b = emu.fit_models(ts, vals, eb, SCREEN, starts, be)
        assert np.array_equal(a["aborted"], b["aborted"]), (label, budget)
        ok = a["aborted"] == 0
        for f in FIELDS:
            bad = np.flatnonzero(ok & (a[f] != b[f]))
            assert len(bad) == 0, f"{label} budget={budget}: {f} differs at starts {starts[bad[:5]]}: {a[bad[:3]]} vs {b[bad[:3]]}"


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_screened_fit_equals_thread_fit(case):
    name, ts, vals, off, ebs = case
    n = len(ts)
    rng = np.random.default_rng(3)
    starts = np.unique(np.concatenate([np.arange(0, min(n, 48)), rng.integers(0, n, 60), np.arange(max(0, n - 20), n)])).astype(np.uint32)
    _assert_same_fits(ts, vals, ebs[0], starts, (None, 40, 300), name)


def _sine(n, base, amp, period, noise, seed, t0, step=1):
    rng = np.random.default_rng(seed)
    i = np.arange(n)
    vals = (base + amp * np.sin(2 * np.pi * i / period + 1.0) + noise * rng.standard_normal(n)).astype(np.float32)
    return (t0 + step * i).astype(np.int64), vals


SINES = [  # base, amplitude, period, noise, first timestamp, interval, bound
    (100.0, 10.0, 1000.0, 0.1, 1_700_000_000_000, 1, (2, 1.0)),      # the benchmark's kind of series, epoch milliseconds
    (60.0, 19.0, 520.0, 0.1, 1_700_000_000_000, 1, (2, 1.0)),
    (140.0, 1.5, 1900.0, 0.1, 1_700_000_000_000, 1, (2, 1.0)),       # models of more than a thousand points
    (100.0, 10.0, 1000.0, 0.1, 0, 1, (2, 1.0)),                       # small timestamps: the reference's lines are accurate
    (100.0, 10.0, 1000.0, 0.1, 1_700_000_000_000_000_000 // 1000, 1000, (2, 5.0)),  # epoch microseconds, 1 ms apart
    (100.0, 10.0, 1000.0, 0.1, -9_000_000_000_000_000, 3600_000, (1, 0.5)),         # near the 2^53 limit, absolute bound
    (0.0, 10.0, 300.0, 0.05, 1_700_000_000_000, 100, (2, 10.0)),      # around zero: relative bound on tiny values
    (1e-30, 1e-31, 700.0, 1e-33, 1_700_000_000_000, 1, (2, 1.0)),     # magnitudes where f32 products underflow
    (100.0, 10.0, 1000.0, 0.0, 1_700_000_000_000, 1, (2, 1.0)),       # no noise: candidates nearly tie
]


@pytest.mark.parametrize("k", range(len(SINES)))
def test_screened_fit_on_sine_series(k):
    """Every fit of the sequential chain (and fits from starts inside its models) on sine + noise series: the data the
    screen is meant for.  Most decisions are certain; the doubtful ones are evaluated exactly -- both paths are exercised."""
    base, amp, period, noise, t0, step, eb = SINES[k]
    n = 30000
    ts, vals = _sine(n, base, amp, period, noise, 7 + k, t0, step)
    starts, cur = [], 0
    while cur < n and len(starts) < 150:  # follow the chain with the one-thread fit
        a = emu.fit_models(ts, vals, eb, 1, np.array([cur], np.uint32), np.array([n], np.uint32))[0]
        starts.append(cur)
        cur = int(a["end"]) + 1 if np.array([a["bpv"]], np.uint32).view(np.float32)[0] <= 4.0 else cur + 1
    rng = np.random.default_rng(k)
    starts = np.unique(np.concatenate([np.array(starts), rng.integers(0, n, 100)])).astype(np.uint32)
    emu.screen_counters()
    _assert_same_fits(ts, vals, eb, starts, (None, 500), f"sine {k}")
    c = emu.screen_counters()
    if k < 4:  # the screen must actually decide these (else the test above says nothing about it)
        assert c["fits"] >= 2 * len(starts) and c["exact_fits"] <= c["fits"] // 10, c


@pytest.mark.parametrize("chunk_len,sched", [(64, (0, 0)), (1000, (5, 3))], ids=["rounds-64", "async-1000"])
@pytest.mark.parametrize("case", [c for c in small_cases() if len(c[1]) <= 8000], ids=lambda c: c[0])
def test_screened_engine_compress_matches_oracle(oracle, case, chunk_len, sched):
    name, ts, vals, off, ebs = case
    want = oracle.compress(ts, vals, off, eb=ebs)
    got = emu.compress(ts, vals, off, eb=ebs, chunk_len=chunk_len, sched_seed=sched[0], in_flight=sched[1], engine=SCREEN)
    assert_segments_equal(got, want, f"{name} chunk_len={chunk_len} sched={sched}")


@pytest.mark.parametrize("seed", range(64))
def test_screened_engine_fuzz(oracle, seed):
    """The fuzz series of the exact engine's test (stitched constants, ramps, noise, repeats, signed zeros, special values;
    random bounds; regular or irregular timestamps) through the screened engine."""
    rng = np.random.default_rng(1000 + seed)
    vals = _fuzz_series(rng)
    n = len(vals)
    step = rng.integers(1, 2000, n) if seed % 4 == 0 else np.full(n, int(rng.integers(1, 5000)))
    ts = (int(rng.integers(0, 2_000_000_000_000_000)) + np.cumsum(step)).astype(np.int64)
    eb = [(0, 0.0), (1, float(10.0 ** rng.integers(-3, 3))), (2, float(rng.choice([0.01, 0.5, 1.0, 5.0, 30.0, 100.0])))][seed % 3 if seed % 7 else 0]
    want = oracle.compress(ts, vals, eb=eb)
    got = emu.compress(ts, vals, eb=eb, chunk_len=int(rng.choice([8, 100, 700])), sched_seed=seed + 1, in_flight=int(rng.choice([1, 2, 9])), engine=SCREEN)
    assert_segments_equal(got, want, f"fuzz seed={seed} eb={eb} n={n}")


@pytest.mark.parametrize("seed", range(24))
def test_screened_engine_fuzz_noisy(oracle, seed):
    """Noisy series with random scale, noise level, trend, bound, epoch and interval: the screen decides nearly everything."""
    rng = np.random.default_rng(5000 + seed)
    n = int(rng.integers(2000, 9000))
    scale = float(10.0 ** rng.integers(-6, 7))
    i = np.arange(n)
    vals = (scale * (rng.normal() * 3 + np.sin(i / rng.uniform(20, 400)) * rng.uniform(0, 2) + i * rng.normal() * 1e-3 +
                     rng.standard_normal(n) * 10.0 ** rng.uniform(-4, -0.5))).astype(np.float32)
    step = int(rng.choice([1, 10, 1000, 60_000, 3_600_000]))
    t0 = int(rng.choice([0, 1_700_000_000, 1_700_000_000_000, 1_700_000_000_000_000, -4_000_000_000_000_000]))
    ts = (t0 + step * i).astype(np.int64)
    eb = (2, float(rng.choice([0.1, 1.0, 5.0, 20.0]))) if seed % 2 else (1, float(scale * rng.choice([0.01, 0.1, 1.0])))
    want = oracle.compress(ts, vals, eb=eb)
    emu.screen_counters()
    got = emu.compress(ts, vals, eb=eb, chunk_len=int(rng.choice([500, 4096])), sched_seed=seed + 1, in_flight=3, engine=SCREEN)
    assert_segments_equal(got, want, f"noisy seed={seed} eb={eb} n={n} scale={scale} t0={t0} step={step}")
    c = emu.screen_counters()
    assert c["fits"] > 0, c


@pytest.mark.parametrize("chunk_len,sched", [(4096, (3, 4)), (700, (9, 2))], ids=["async-4096", "async-700"])
def test_screened_engine_on_long_models(oracle, chunk_len, sched):
    for name, ts, vals, eb in long_model_cases():
        want = oracle.compress(ts, vals, eb=eb)
        got = emu.compress(ts, vals, eb=eb, chunk_len=chunk_len, sched_seed=sched[0], in_flight=sched[1], engine=SCREEN)
        assert_segments_equal(got, want, f"{name} chunk_len={chunk_len}")
