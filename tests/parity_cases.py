"""Seeded workloads and comparison helpers shared by the CPU (kernel-body emulation) and GPU parity tests.

Every case is (name, timestamps, values, unit_off, error_bounds).  They cover the shapes the reference's
tests use (compression.rs:421-863: constant / linear / random, regular / irregular, lossless / abs / rel)
plus the five BASELINE.json configs at sizes the oracle finishes in seconds, plus the edge cases:
empty and 1/2/3-point units, ragged units, NaN / inf / signed zeros, > 255 residual runs, all models.
"""
import numpy as np

from modelardb_rs_b200 import synthetic as syn

LOSSLESS = (0, 0.0)


def _cat(units):
    ts = np.concatenate([u[0] for u in units]) if units else np.zeros(0, np.int64)
    vals = np.concatenate([u[1] for u in units]) if units else np.zeros(0, np.float32)
    off = np.zeros(len(units) + 1, np.uint64)
    off[1:] = np.cumsum([len(u[0]) for u in units])
    return ts.astype(np.int64), vals.astype(np.float32), off


def small_cases():
    cases = []
    for irregular in (False, True):
        for noise in (None, (1.0, 1.05)):
            for eb in (LOSSLESS, (1, 5.0), (2, 5.0)):
                ts, v = syn.mixed_series(6000, 31 + int(irregular), irregular=irregular, noise=noise)
                cases.append((f"mixed-irr{int(irregular)}-noise{noise is not None}-eb{eb}", ts, v,
                              np.array([0, len(ts)], np.uint64), [eb]))
    # epoch-scale timestamps: cancellation-sensitive Swing arithmetic (quirk Q1)
    for eb in (LOSSLESS, (2, 1.0), (2, 10.0), (1, 0.5)):
        ts, v, off = syn.multi_series(6, 5000, 2, "sine")
        cases.append((f"sine-epoch-eb{eb}", ts, v, off, [eb] * 6))
        ts, v, off = syn.multi_series(6, 5000, 3, "walk")
        cases.append((f"walk-epoch-eb{eb}", ts, v, off, [eb] * 6))
    ts, v, off = syn.multi_series(4, 4000, 5, "walk", irregular=True)
    cases.append(("walk-irregular-rel1", ts, v, off, [(2, 1.0)] * 4))
    ts, v, off = syn.multi_series(4, 4000, 5, "sine", irregular=True)
    cases.append(("sine-irregular-rel1", ts, v, off, [(2, 1.0)] * 4))
    # ragged / tiny / empty units, one bound per unit
    units, ebs = [], []
    rng = np.random.default_rng(9)
    for k, n in enumerate([0, 1, 2, 3, 7, 8, 9, 15, 16, 17, 0, 263, 264, 265, 1000, 1, 0]):
        t = syn.regular_timestamps(n) if k % 2 == 0 else syn.irregular_timestamps(max(n, 1), 100 + k)[:n]
        if k % 3 == 0:
            v = syn.sine_noise(n, 200 + k)
        elif k % 3 == 1:
            v = rng.uniform(-1e3, 1e3, n).astype(np.float32)
        else:
            v = np.full(n, 42.5, np.float32)
        units.append((t, v))
        ebs.append([LOSSLESS, (1, 1.0), (2, 5.0)][k % 3])
    ts, v, off = _cat(units)
    cases.append(("ragged-units", ts, v, off, ebs))
    # special values
    sp = np.array([0.0, -0.0, 0.0, -0.0, np.nan, np.nan, np.inf, np.inf, -np.inf, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0,
                   1.0, np.nan, 3.0, np.float32(1e-45), np.float32(-1e-45), 3.4e38, -3.4e38] * 20, np.float32)
    for eb in (LOSSLESS, (1, 1.0), (2, 10.0)):
        cases.append((f"specials-eb{eb}", syn.regular_timestamps(len(sp)), sp, np.array([0, len(sp)], np.uint64), [eb]))
    # runs long enough for a model of NaN / inf / -0.0 to be stored (PMC-Mean and Swing special paths)
    runs = np.concatenate([np.full(12, np.nan), np.full(9, np.inf), np.full(10, -np.inf), np.full(8, -0.0),
                           np.full(8, 0.0), [1.0, 2.0, 3.0], np.full(20, np.nan), [5.0], np.full(8, np.inf)]).astype(np.float32)
    for eb in (LOSSLESS, (1, 1.0), (2, 10.0)):
        cases.append((f"special-runs-eb{eb}", syn.regular_timestamps(len(runs)), runs, np.array([0, len(runs)], np.uint64), [eb]))
    # a model followed by > 255 incompressible points -> separate MacaqueV row (compression.rs:329-349)
    v = np.concatenate([np.full(50, 7.0, np.float32), rng.uniform(-1e6, 1e6, 700).astype(np.float32),
                        np.full(50, 9.0, np.float32), rng.uniform(-1e6, 1e6, 255).astype(np.float32),
                        np.full(20, 1.0, np.float32), rng.uniform(-1e6, 1e6, 256).astype(np.float32)])
    for eb in (LOSSLESS, (2, 1.0)):
        cases.append((f"long-residual-runs-eb{eb}", syn.regular_timestamps(len(v)), v, np.array([0, len(v)], np.uint64), [eb]))
    # lossy MacaqueV with tiny / huge bounds (rewrite position wraps, macaque_v.rs:333-336)
    v = rng.uniform(-1e3, 1e3, 600).astype(np.float32)
    for eb in ((1, 1e-10), (1, 1e10), (2, 1e-6), (2, 100.0), (1, 3.0e38)):
        cases.append((f"lossy-extreme-eb{eb}", syn.regular_timestamps(len(v)), v, np.array([0, len(v)], np.uint64), [eb]))
    return cases


def assert_segments_equal(a, b, where=""):
    """Bit-exact comparison of two segment batches (oracle.Segments-like objects)."""
    assert len(a) == len(b), f"{where}: {len(a)} vs {len(b)} segments"
    for col in ("model_type_id", "start_time", "end_time", "timestamps_off", "values_off", "residuals_off",
                "timestamps_data", "values_data", "residuals_data"):
        x, y = getattr(a, col), getattr(b, col)
        if not np.array_equal(x, y):
            bad = np.flatnonzero(x != y)[:5] if len(x) == len(y) else "length"
            raise AssertionError(f"{where}: column {col} differs at {bad}")
    for col in ("min_value", "max_value"):
        x, y = getattr(a, col).view(np.uint32), getattr(b, col).view(np.uint32)
        if not np.array_equal(x, y):
            bad = np.flatnonzero(x != y)[:5]
            raise AssertionError(f"{where}: column {col} differs (bit pattern) at rows {bad}: "
                                 f"{getattr(a, col)[bad]} vs {getattr(b, col)[bad]}")
    if a.unit_seg_off is not None and b.unit_seg_off is not None:
        assert np.array_equal(a.unit_seg_off, b.unit_seg_off), f"{where}: unit_seg_off differs"


def assert_f32_bits_equal(x, y, where="", nan_payload_matters=True):
    """Bit-pattern equality.  nan_payload_matters=False is for values PRODUCED BY ARITHMETIC (sums): there
    the sign/payload of a NaN is unspecified in Rust and hardware-dependent, so NaN == NaN."""
    x, y = np.asarray(x, np.float32).copy(), np.asarray(y, np.float32).copy()
    if not nan_payload_matters:
        x[np.isnan(x)] = np.float32(np.nan)
        y[np.isnan(y)] = np.float32(np.nan)
    x, y = x.view(np.uint32), y.view(np.uint32)
    assert len(x) == len(y), f"{where}: length {len(x)} vs {len(y)}"
    if not np.array_equal(x, y):
        bad = np.flatnonzero(x != y)
        raise AssertionError(f"{where}: {len(bad)} of {len(x)} values differ, first at {bad[:5]}: "
                             f"{x.view(np.float32)[bad[:5]]} vs {y.view(np.float32)[bad[:5]]}")


def long_model_cases():
    """Series of 20 000 points with models of thousands of points, values right around the bounds, zeros, sign changes,
    special values inside, extreme magnitudes: (name, timestamps, values, error bound)."""
    n = 20_000
    i = np.arange(n)
    rng = np.random.default_rng(42)
    ts = (1_600_000_000_000_000 + 1000 * i).astype(np.int64)
    noise = rng.standard_normal(n)
    cases = []

    def add(name, vals, ebs, t=ts):
        for eb in ebs:
            cases.append((f"{name}-eb{eb[0]}:{eb[1]}", t, np.asarray(vals, np.float32), eb))

    add("constant", np.full(n, 100.0), [(0, 0.0), (1, 0.5), (2, 1.0)])
    add("constant-noise", 100.0 + 0.1 * noise, [(1, 1.0), (2, 1.0), (2, 0.2)])
    # PMC-Mean's relative test at assorted bounds, values scattered right around each bound
    for pct in (1e-4, 0.37, 3.0, 33.3, 100.0):
        add(f"around-{pct}pct", 250.0 * (1.0 + (pct / 100.0) * 0.9 * np.clip(noise, -1.3, 1.3)), [(2, pct)])
        add(f"negative-around-{pct}pct", -0.004 * (1.0 + (pct / 100.0) * 0.9 * np.clip(noise, -1.3, 1.3)), [(2, pct)])
    add("slow-ramp", 100.0 + 0.001 * i + 0.01 * noise, [(1, 0.1), (2, 0.05), (2, 1.0)])
    add("steps", np.where(i < 9_000, 100.0, 150.0) + 0.05 * noise, [(1, 1.0), (2, 1.0)])
    add("zeros", np.zeros(n), [(0, 0.0), (1, 0.5), (2, 1.0)])
    add("signed-zeros", np.where(i % 3 == 0, -0.0, 0.0), [(0, 0.0), (1, 0.5), (2, 1.0)])
    add("zero-touching", np.maximum(0.0, 0.5 * np.sin(i / 900.0)), [(1, 1.0), (2, 10.0)])
    add("sign-change", -5.0 + 10.0 * i / n, [(1, 10.0), (1, 0.01), (2, 5.0)])
    t_irr = ts.copy()
    t_irr[7_000:] += 137
    add("late-irregular", 100.0 + 0.05 * noise, [(1, 1.0), (2, 1.0)], t_irr)
    v = 100.0 + 0.05 * noise
    v_nan = v.copy(); v_nan[9_000] = np.nan
    v_inf = v.copy(); v_inf[9_001] = np.inf
    add("nan-inside", v_nan, [(1, 1.0), (2, 1.0)])
    add("inf-inside", v_inf, [(1, 1.0), (2, 1.0)])
    add("huge", np.full(n, 1e30) * (1.0 + 1e-4 * noise), [(2, 1.0), (1, 1e28)])
    add("tiny", np.full(n, 1e-35) * (1.0 + 1e-3 * noise), [(2, 1.0), (1, 1e-36)])
    add("subnormal", np.full(n, 1e-41) * (1.0 + 1e-2 * noise), [(2, 5.0), (1, 1e-42)])
    add("wide-exponents", np.where(i % 2 == 0, 1e10, 1e-10), [(1, 1e11)])
    for hi in (101.9, 102.0, 102.02, 102.05, 102.5):  # PMC-Mean close to a 1 % relative bound from either side
        add(f"alternating-{hi}", np.where(i % 2 == 0, 100.0, hi), [(2, 1.0)])
    for amp in (0.45, 0.5, 0.55):  # Swing close to an absolute bound of 0.5 around a line
        add(f"sawtooth-{amp}", 10.0 + 0.002 * i + amp * np.where(i % 2 == 0, 1.0, -1.0), [(1, 0.5)])
    add("plateau-then-noise", np.where(i < 12_345, 42.0, 42.0 + 5.0 * noise), [(0, 0.0), (2, 1.0)])
    return cases
