"""Randomised parity soak on a GPU: batches of series with random scale, trend, noise, quantisation, error bound, epoch and
sampling interval through mdbcu_compress (automatic engine: the screened fit) and through the oracle on the host cores; every
column of every segment must be bit-identical, and grid / aggregate of the result must match too.

usage: python tools/gpu_soak.py [seconds] [first seed]      (prints one JSON line; exit code 1 on the first difference)
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from modelardb_rs_b200 import compression as mc  # noqa: E402
from oracle import mdb_oracle as oracle  # noqa: E402
from tests.parity_cases import assert_f32_bits_equal, assert_segments_equal  # noqa: E402


def batch(seed):
    rng = np.random.default_rng(seed)
    n_series = int(rng.integers(8, 400))
    n = int(rng.integers(2_000, 60_000))
    step = int(rng.choice([1, 10, 1000, 60_000, 3_600_000]))
    t0 = int(rng.choice([0, 1_700_000_000, 1_700_000_000_000, 1_700_000_000_000_000, -4_000_000_000_000_000, 8_000_000_000_000_000 - n * step]))
    i = np.arange(n)
    vals = np.empty((n_series, n), np.float32)
    for s in range(n_series):
        scale = float(10.0 ** rng.integers(-6, 7))
        v = scale * (rng.normal() * 3 + np.sin(i / rng.uniform(20, 800) + rng.uniform(0, 6)) * rng.uniform(0, 2) + i * rng.normal() * 1e-4 +
                     rng.standard_normal(n) * 10.0 ** rng.uniform(-5, -0.5))
        kind = rng.integers(0, 10)
        if kind == 0:
            v = np.round(v / scale, int(rng.integers(0, 3))) * scale  # quantised: ties everywhere
        elif kind == 1:
            v[rng.integers(0, n, 5)] = rng.choice([np.nan, np.inf, -np.inf, 0.0, -0.0, 3.4e38])  # special values inside
        elif kind == 2:
            v = np.full(n, v[0])  # constant
        elif kind == 3:
            v = v[0] + scale * 1e-3 * i  # a ramp
        vals[s] = v.astype(np.float32)
    ts = np.tile((t0 + step * i).astype(np.int64), n_series)
    if seed % 7 == 3:  # one irregular unit among the regular ones
        ts[:n] = t0 + np.cumsum(rng.integers(1, 2 * step + 2, n))
    off = (np.arange(n_series + 1, dtype=np.uint64) * np.uint64(n)).astype(np.uint64)
    ebs = []
    for s in range(n_series):
        k = int(rng.integers(0, 10))
        ref = float(abs(vals[s, 0])) if np.isfinite(vals[s, 0]) and abs(vals[s, 0]) < 1e30 else 1.0
        ebs.append((0, 0.0) if k == 0 else (1, float(ref * rng.choice([0.001, 0.01, 0.1]) + 1e-30)) if k <= 3
                   else (2, float(rng.choice([0.01, 0.1, 1.0, 5.0, 20.0, 100.0]))))
    return ts, vals.reshape(-1), off, ebs


def main():
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    ctx = mc.Context(0)
    t_end = time.time() + seconds
    points = batches = segments = 0
    threads = os.cpu_count() or 8
    while time.time() < t_end:
        ts, vals, off, ebs = batch(seed)
        want = oracle.compress(ts, vals, off, eb=ebs, n_threads=threads)
        seg = mc.compress(ts, vals, off, [mc.ErrorBound(*e) for e in ebs], ctx)
        got = seg.to_host()
        try:
            assert_segments_equal(got, want, f"seed {seed}")
            wts, wval, _ = oracle.grid(want, n_threads=threads)
            gts, gval = mc.grid(seg, ctx=ctx)
            assert np.array_equal(gts.cpu().numpy(), wts), f"seed {seed}: grid timestamps"
            assert_f32_bits_equal(gval.cpu().numpy(), wval, f"seed {seed}: grid values")
            c, mn, mx, sm = mc.aggregate(got, want.unit_seg_off, ctx)
            wc, wmn, wmx, wsm = oracle.aggregate(want, want.unit_seg_off)
            assert np.array_equal(c, wc), f"seed {seed}: counts"
            # (values, not bit patterns: which of +0.0 / -0.0 a fold over both returns is unspecified in the reference itself,
            # f32::min / f32::max on a tie -- DESIGN.md section 2, unpinned corner 1)
            assert np.array_equal(mn, wmn, equal_nan=True), f"seed {seed}: min"
            assert np.array_equal(mx, wmx, equal_nan=True), f"seed {seed}: max"
            ok = np.isclose(sm, wsm, rtol=1e-12, atol=0.0) | (np.isnan(sm) & np.isnan(wsm)) | (sm == wsm)
            assert ok.all(), f"seed {seed}: sums"
        except AssertionError as e:
            print(json.dumps({"soak": "FAILED", "seed": seed, "error": str(e)[:500], "batches_ok": batches}))
            sys.exit(1)
        finally:
            seg.free()
        points += len(ts)
        segments += len(want)
        batches += 1
        seed += 1
    print(json.dumps({"soak": "ok", "batches": batches, "points": points, "segments": segments, "next_seed": seed, "seconds": seconds}))


if __name__ == "__main__":
    main()
