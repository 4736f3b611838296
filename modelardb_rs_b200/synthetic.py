"""Seeded synthetic time series for the tests and the benchmark.

The shapes follow the reference's test generators
(crates/modelardb_test/src/data_generation.rs:108-284: constant, linear, uniform random, optional
noise; regular 100-step or irregular U[100,200) timestamps) and the five BASELINE.json configs.
The Rust `StdRng` stream is not reproducible outside Rust, so these are numpy `default_rng` streams
with the same structure.  Epoch-scale microsecond timestamps are used on purpose: they make Swing's
f64 `slope * t + intercept` cancellation-sensitive, so an FMA contraction or a reordered operation on
the GPU shows up as a bit mismatch instead of hiding.
"""
from __future__ import annotations

import numpy as np

EPOCH_US = 1_600_000_000_000_000  # 2020-09-13 in microseconds
STEP_US = 1000                    # 1 ms sampling


def regular_timestamps(n: int, start: int = EPOCH_US, step: int = STEP_US) -> np.ndarray:
    return start + step * np.arange(n, dtype=np.int64)


def irregular_timestamps(n: int, seed: int, start: int = EPOCH_US, lo: int = 100, hi: int = 200) -> np.ndarray:
    """data_generation.rs:208-221: t += U[100, 200)."""
    rng = np.random.default_rng(seed)
    d = rng.integers(lo, hi, size=n, dtype=np.int64)
    d[0] = 0
    return start + np.cumsum(d)


def sine_noise(n: int, seed: int, base=100.0, amp=10.0, period=1000.0, sigma=0.1, phase=0.0) -> np.ndarray:
    """cfg1/cfg2/cfg4/cfg5 generator: base + amp*sin(2*pi*i/period + phase) + N(0, sigma) -> f32."""
    rng = np.random.default_rng(seed)
    i = np.arange(n, dtype=np.float64)
    return (base + amp * np.sin(2.0 * np.pi * i / period + phase) + rng.normal(0.0, sigma, n)).astype(np.float32)


def random_walk(n: int, seed: int, base=100.0, sigma=1.0) -> np.ndarray:
    """cfg3 generator: base + cumsum(N(0, sigma)) -> f32 (high entropy, MacaqueV-heavy when lossless)."""
    rng = np.random.default_rng(seed)
    return (base + np.cumsum(rng.normal(0.0, sigma, n))).astype(np.float32)


def constant(n: int, seed: int, noise=None) -> np.ndarray:
    rng = np.random.default_rng(seed)
    v = np.full(n, rng.random(dtype=np.float32), dtype=np.float32)
    if noise is not None:
        v = v + rng.uniform(noise[0], noise[1], n).astype(np.float32)
    return v.astype(np.float32)


def linear(ts: np.ndarray, seed: int, noise=None) -> np.ndarray:
    """data_generation.rs:244-257: (slope * t + intercept) as f32 with integer slope in [-10,10) \\ {0}."""
    rng = np.random.default_rng(seed)
    slope = 0
    while slope == 0:
        slope = int(rng.integers(-10, 10))
    intercept = int(rng.integers(1, 50))
    v = (slope * ts.astype(np.int64) + intercept).astype(np.float32)
    if noise is not None:
        v = v + rng.uniform(noise[0], noise[1], len(ts)).astype(np.float32)
    return v.astype(np.float32)


def uniform_random(n: int, seed: int, lo=0.0, hi=100.0) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return rng.uniform(lo, hi, n).astype(np.float32)


def mixed_series(n: int, seed: int, irregular: bool = False, noise=None, seg_len=(50, 501),
                 random_range=(100.0, 200.0), small_timestamps: bool = True):
    """generate_univariate_time_series (data_generation.rs:108-203): runs of constant / linear /
    random values of random length.  Returns (timestamps, values)."""
    rng = np.random.default_rng(seed)
    if small_timestamps:  # the reference's own tests use 0, 100, 200, ...
        ts = irregular_timestamps(n, seed + 1, start=0) if irregular else regular_timestamps(n, 0, 100)
    else:
        ts = irregular_timestamps(n, seed + 1) if irregular else regular_timestamps(n)
    out = np.empty(n, dtype=np.float32)
    pos = 0
    k = 0
    while pos < n:
        ln = int(rng.integers(seg_len[0], seg_len[1]))
        end = min(n, pos + ln)
        kind = int(rng.integers(0, 3))
        if kind == 0:
            out[pos:end] = constant(end - pos, seed * 7919 + k, noise)
        elif kind == 1:
            out[pos:end] = linear(ts[pos:end] - (0 if small_timestamps else EPOCH_US), seed * 7919 + k, noise)
        else:
            out[pos:end] = uniform_random(end - pos, seed * 7919 + k, *random_range)
        pos = end
        k += 1
    return ts, out


def multi_series(n_series: int, n_points: int, seed: int, kind: str = "sine", irregular: bool = False):
    """Concatenated units: returns (timestamps[n_series*n_points], values[...], unit_off[n_series+1])."""
    ts = np.empty(n_series * n_points, dtype=np.int64)
    vals = np.empty(n_series * n_points, dtype=np.float32)
    rng = np.random.default_rng(seed)
    for s in range(n_series):
        sl = slice(s * n_points, (s + 1) * n_points)
        ts[sl] = irregular_timestamps(n_points, seed + 1000 + s) if irregular else regular_timestamps(n_points)
        if kind == "sine":
            vals[sl] = sine_noise(n_points, seed + s, base=float(rng.uniform(50, 150)), amp=float(rng.uniform(1, 20)),
                                  period=float(rng.uniform(500, 2000)), phase=float(rng.uniform(0, 6.28)))
        elif kind == "walk":
            vals[sl] = random_walk(n_points, seed + s)
        elif kind == "mixed":
            vals[sl] = mixed_series(n_points, seed + s, irregular=False, noise=(1.0, 1.05))[1]
        else:
            raise ValueError(kind)
    unit_off = (np.arange(n_series + 1, dtype=np.uint64) * np.uint64(n_points)).astype(np.uint64)
    return ts, vals, unit_off
