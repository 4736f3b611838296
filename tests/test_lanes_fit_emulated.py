"""The one-lane-per-chain engine (modelardb_rs_b200/csrc/mdb_fit_lanes.cuh) run on the HOST.  A debugging harness for the
GPU-less build container: the lane's state machine (LaneChain: PMC-Mean and Swing point by point on a regular unit,
timestamps derived from the index, Swing's error sums deferred) steps every chunk from its first index, exactly as
k_spec_lanes does, and the asynchronous scheduler then stitches the chunks with the warp-cooperative fit -- the product's
default compress path.  Everything is compared with the oracle, column for column; the same comparisons run on the
device through the C-ABI in the GPU tests (whose default engine is this one)."""
import numpy as np
import pytest

from tests import emu_lib as emu
from tests.parity_cases import assert_segments_equal, long_model_cases, small_cases
from tests.test_warp_fit_emulated import _fuzz_series


@pytest.mark.parametrize("chunk_len,warmup,sched", [(64, 0, (2, 1)), (64, 100, (3, 2)), (1000, 300, (5, 3)), (1000, 5000, (6, 2)), (4096, 0, (7, 8))],
                         ids=["64", "64+100", "1000+300", "1000+5000", "4096"])
@pytest.mark.parametrize("case", [c for c in small_cases() if len(c[1]) <= 30000], ids=lambda c: c[0])
def test_emulated_lanes_compress_matches_oracle(oracle, case, chunk_len, warmup, sched):
    name, ts, vals, off, ebs = case
    want = oracle.compress(ts, vals, off, eb=ebs)
    got = emu.compress(ts, vals, off, eb=ebs, chunk_len=chunk_len, sched_seed=sched[0], in_flight=sched[1], engine=2, lanes=True, lane_warmup=warmup,
                       lane_rounds=sched[0] % 5)
    assert_segments_equal(got, want, f"{name} chunk_len={chunk_len} sched={sched}")
    assert emu.division_mismatches() == 0


def test_lanes_run_regular_units_and_leave_the_others(oracle):
    """Regular units are walked by lanes; units with an irregular interval anywhere, with fewer than two points or with
    timestamps of 2^53 and beyond are not; a chunk in which a lane meets a NaN or an infinity is abandoned and run by the
    cooperative engine.  The result is the oracle's either way."""
    n = 5000
    rng = np.random.default_rng(4)
    vals = (50 + np.cumsum(rng.normal(0, 0.05, n))).astype(np.float32)
    regular = 1_600_000_000_000_000 + 1000 * np.arange(n, dtype=np.int64)
    cases = {
        "regular": (regular, vals, True, 0),
        "one odd interval at the end": (np.concatenate([regular[:-1], regular[-1:] + 1]), vals, False, 0),
        "one odd interval in the middle": (np.concatenate([regular[:2500], regular[2500:] + 7]), vals, False, 0),
        "beyond 2^53": (regular + (1 << 53), vals, False, 0),
        "negative, regular": (regular - 3_000_000_000_000_000, vals, True, 0),
        "a NaN in one chunk": (regular, np.concatenate([vals[:3000], [np.float32(np.nan)], vals[3001:]]).astype(np.float32), True, 1),
    }
    for name, (ts, v, runs, bails_min) in cases.items():
        for eb in ((2, 0.5), (0, 0.0), (1, 0.05)):
            want = oracle.compress(ts, v, eb=eb)
            before = emu.lane_counters()
            got = emu.compress(ts, v, eb=eb, chunk_len=512, sched_seed=11, in_flight=3, engine=2, lanes=True, lane_warmup=200)
            after = emu.lane_counters()
            assert_segments_equal(got, want, f"{name} {eb}")
            assert (after[0] > before[0]) == runs, name
            assert after[1] - before[1] >= bails_min, name


@pytest.mark.parametrize("seed", range(48))
def test_emulated_lanes_fuzz(oracle, seed):
    """Random stitched series (constants, ramps, noise, zeros of both signs, special values), random bounds, regular
    timestamps of random phase and step, random chunk lengths and schedules."""
    rng = np.random.default_rng(7000 + seed)
    vals = np.concatenate([_fuzz_series(rng) for _ in range(3)])
    n = len(vals)
    ts = (int(rng.integers(-10**15, 2 * 10**15)) + int(rng.integers(1, 100000)) * np.arange(n, dtype=np.int64)).astype(np.int64)
    eb = [(0, 0.0), (1, float(10.0 ** rng.integers(-3, 3))), (2, float(rng.choice([0.01, 0.5, 1.0, 5.0, 30.0, 100.0])))][seed % 3]
    want = oracle.compress(ts, vals, eb=eb)
    got = emu.compress(ts, vals, eb=eb, chunk_len=int(rng.choice([8, 24, 100, 700])), sched_seed=seed + 1, in_flight=int(rng.choice([1, 2, 9])),
                       engine=2, lanes=True, lane_warmup=int(rng.choice([0, 5, 60, 1000])), lane_rounds=int(rng.integers(0, 5)))
    assert_segments_equal(got, want, f"fuzz seed={seed} eb={eb} n={n}")


def test_emulated_lanes_on_long_models(oracle):
    """Models of thousands of points (a lane's fit is cut one chunk length past its chunk and resumed by the cooperative
    engine), values right around the bounds, zeros, sign changes, extreme magnitudes."""
    for name, ts, vals, eb in long_model_cases():
        want = oracle.compress(ts, vals, eb=eb)
        for chunk_len in (700, 4096):
            got = emu.compress(ts, vals, eb=eb, chunk_len=chunk_len, sched_seed=3, in_flight=4, engine=2, lanes=True, lane_warmup=chunk_len // 2)
            assert_segments_equal(got, want, f"{name} chunk_len={chunk_len}")
    assert emu.division_mismatches() == 0
