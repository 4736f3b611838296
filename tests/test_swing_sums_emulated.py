"""The body of k_swing_finish (csrc/mdb_swing_sums.cuh: a lane adds the two error sums of one Swing model in point order, reading
quads of values two ahead through three buffers) run on the host against the plain loop of swing_finish: every alignment of
the first point, every length around the quad loop's corners, values equal to the first one (their terms are skipped), special
values, regular and irregular units.  Bit-identical sums -- the GPU parity tests compare the finished models."""
import ctypes as C

import numpy as np
import pytest

from tests import emu_lib as emu


def _sums(ts, vals, regular, start, end):
    out = (C.c_double * 4)()
    emu.lib().emu_swing_sums(ts.ctypes.data_as(C.c_void_p), vals.ctypes.data_as(C.c_void_p), int(regular), C.c_uint32(start), C.c_uint32(end), out)
    return np.array(out[:2]).view(np.uint64), np.array(out[2:]).view(np.uint64)


def _series(rng, n, kind):
    i = np.arange(n)
    v = (100.0 + 10.0 * np.sin(i / 40.0) + rng.standard_normal(n)).astype(np.float32)
    if kind == "repeats":
        v = np.round(v).astype(np.float32)  # many points equal to a model's first value
    elif kind == "special":
        v[rng.integers(0, n, n // 50)] = rng.choice(np.array([np.nan, np.inf, -np.inf, 0.0, -0.0, 3.4e38], np.float32), n // 50)
    elif kind == "tiny":
        v = (v * 1e-38).astype(np.float32)
    return v


@pytest.mark.parametrize("kind", ["noise", "repeats", "special", "tiny"])
@pytest.mark.parametrize("step", [1, 1000, 3_600_000])
def test_lane_sums_equal_the_plain_loop(kind, step):
    rng = np.random.default_rng(["noise", "repeats", "special", "tiny"].index(kind) * 10 + step % 7)
    n = 3000
    vals = _series(rng, n, kind)
    ts = (1_700_000_000_000 + step * np.arange(n)).astype(np.int64)
    cases = [(s, s + k) for s in range(0, 9) for k in range(0, 40)]  # every alignment x every short length
    cases += [(int(s), int(s + k)) for s, k in zip(rng.integers(0, 1000, 300), rng.integers(1, 1900, 300))]
    for start, end in cases:
        got, want = _sums(ts, vals, True, start, end)
        assert np.array_equal(got, want), (kind, step, start, end)


def test_irregular_unit_reads_the_timestamps():
    rng = np.random.default_rng(5)
    n = 2000
    vals = _series(rng, n, "repeats")
    ts = (1_700_000_000_000 + np.cumsum(rng.integers(1, 5000, n))).astype(np.int64)
    for start, k in zip(rng.integers(0, 1000, 200), rng.integers(1, 900, 200)):
        got, want = _sums(ts, vals, False, int(start), int(start + k))
        assert np.array_equal(got, want), (start, k)
