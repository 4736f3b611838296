"""Randomised parity soak WITHOUT a GPU: the series of tools/gpu_soak.py (smaller batches) through the screened fit engine run on
the host by the lane emulator (tests/emu: the same csrc/ headers, 32 fibers per warp, the asynchronous chunk scheduler stepped
in a seeded random order) and through the oracle; every column of every segment must be bit-identical.

usage: python tools/emu_soak.py [seconds] [first seed] [seed stride] [smooth]     (one JSON line; exit code 1 on the first difference)
Run several with different first seeds and the same stride to use several cores.  `smooth`: hardly any noise and loose bounds,
i.e. models of thousands of points (quiet steps, long pending Swing models, fits that outgrow their chunk).
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import mdb_oracle as oracle  # noqa: E402
from tests import emu_lib as emu  # noqa: E402
from tests.parity_cases import assert_segments_equal  # noqa: E402


def series(seed, smooth=False):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(500, 12_000))
    step = int(rng.choice([1, 10, 1000, 60_000, 3_600_000]))
    t0 = int(rng.choice([0, 1_700_000_000, 1_700_000_000_000, 1_700_000_000_000_000, -4_000_000_000_000_000, 8_000_000_000_000_000 - n * step]))
    i = np.arange(n)
    scale = float(10.0 ** rng.integers(-6, 7))
    v = scale * (rng.normal() * 3 + np.sin(i / rng.uniform(20, 800) + rng.uniform(0, 6)) * rng.uniform(0, 2) + i * rng.normal() * 1e-4 +
                 rng.standard_normal(n) * 10.0 ** (rng.uniform(-8, -3.5) if smooth else rng.uniform(-5, -0.5)))
    kind = rng.integers(0, 10)
    if kind == 0:
        v = np.round(v / scale, int(rng.integers(0, 3))) * scale  # quantised: ties everywhere
    elif kind == 1:
        v[rng.integers(0, n, 5)] = rng.choice([np.nan, np.inf, -np.inf, 0.0, -0.0, 3.4e38])  # special values inside
    elif kind == 2:
        v = np.full(n, v[0])  # constant
    elif kind == 3:
        v = v[0] + scale * 1e-3 * i  # a ramp
    vals = v.astype(np.float32)
    ts = (t0 + step * i).astype(np.int64)
    if seed % 7 == 3:
        ts = (t0 + np.cumsum(rng.integers(1, 2 * step + 2, n))).astype(np.int64)
    k = int(rng.integers(4 if smooth else 0, 10))
    ref = float(abs(vals[0])) if np.isfinite(vals[0]) and abs(vals[0]) < 1e30 else 1.0
    eb = (0, 0.0) if k == 0 else (1, float(ref * rng.choice([0.001, 0.01, 0.1]) + 1e-30)) if k <= 3 else (2, float(rng.choice([5.0, 20.0, 50.0, 100.0] if smooth else [0.01, 0.1, 1.0, 5.0, 20.0, 100.0])))
    chunk_len = int(rng.choice([64, 700, 4096]))
    return ts, vals, eb, chunk_len, int(rng.choice([1, 2, 5]))


def main():
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    stride = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    smooth = len(sys.argv) > 4 and sys.argv[4] == "smooth"
    t_end = time.time() + seconds
    points = runs = 0
    while time.time() < t_end:
        ts, vals, eb, chunk_len, in_flight = series(seed, smooth)
        want = oracle.compress(ts, vals, eb=eb)
        got = emu.compress(ts, vals, eb=eb, chunk_len=chunk_len, sched_seed=seed + 1, in_flight=in_flight, engine=5)
        try:
            assert_segments_equal(got, want, f"seed {seed}")
        except AssertionError as e:
            print(json.dumps({"soak": "FAILED", "seed": seed, "eb": eb, "chunk_len": chunk_len, "error": str(e)[:400], "runs_ok": runs}))
            sys.exit(1)
        points += len(ts)
        runs += 1
        seed += stride
    print(json.dumps({"soak": "ok", "series": runs, "points": points, "next_seed": seed, "stride": stride, "smooth": smooth, "seconds": seconds, "counters": emu.screen_counters()}))


if __name__ == "__main__":
    main()
