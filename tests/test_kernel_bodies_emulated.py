"""CPU check of the per-thread kernel bodies (csrc/mdb_*.cuh compiled for the host, tests/emu/emu.cc)
against the oracle.  This is the fast inner loop on the GPU-less build box; the bit-exactness CLAIM is
made by the `-m gpu` tests, which run the real kernels through the C-ABI on the same cases."""
import numpy as np
import pytest

from tests import emu_lib as emu
from tests.parity_cases import assert_f32_bits_equal, assert_segments_equal, small_cases

CASES = small_cases()


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_compress_grid_aggregate_bodies_match_oracle(oracle, case):
    name, ts, vals, off, ebs = case
    want = oracle.compress(ts, vals, off, eb=ebs)
    got = emu.compress(ts, vals, off, eb=ebs)
    assert_segments_equal(got, want, name)

    wts, wval, woff = oracle.grid(want)
    gts, gval, goff = emu.grid(want)
    assert np.array_equal(goff, woff)
    assert np.array_equal(gts, wts)
    assert_f32_bits_equal(gval, wval, name + " grid")
    assert np.array_equal(gts, ts)  # timestamps always round-trip exactly (compression.rs:912)

    wsum = oracle.segment_sums(want)
    gsum, gcount = emu.segment_sums(want)
    assert_f32_bits_equal(gsum, wsum, name + " segment sums", nan_payload_matters=False)
    assert np.array_equal(gcount, np.diff(woff))

    wc, wmn, wmx, wsm = oracle.aggregate(want, want.unit_seg_off)
    gc, gmn, gmx, gsm = emu.aggregate(want, want.unit_seg_off)
    assert np.array_equal(gc, wc)
    assert_f32_bits_equal(gmn, wmn, name + " min")
    assert_f32_bits_equal(gmx, wmx, name + " max")
    assert np.array_equal(gsm.view(np.uint64), wsm.view(np.uint64))  # same fold order on the host
