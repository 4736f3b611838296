"""The committed fixtures of tests/golden/ (made by tests/golden/make_golden.py from the CPU oracle, which
test_oracle_golden.py pins against the reference's own vectors): the oracle must keep reproducing them
(`not gpu`), and the CUDA path must reproduce them through the C-ABI (`gpu`)."""
import glob
import hashlib
import os

import numpy as np
import pytest

from tests.golden.make_golden import SEGMENT_COLUMNS
from tests.parity_cases import assert_f32_bits_equal

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = sorted(glob.glob(os.path.join(HERE, "golden", "*.npz")))


def _sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), np.uint8)


def _check(fix, seg, gts, gval, agg, where, sum_exact):
    for c in SEGMENT_COLUMNS:
        want, got = fix["seg_" + c], np.asarray(getattr(seg, c))
        assert want.dtype == got.dtype and want.shape == got.shape, (where, c, want.shape, got.shape)
        assert np.array_equal(want.view(np.uint8), got.view(np.uint8)), (where, c)  # bit-exact, NaN payloads included
    assert len(gts) == int(fix["grid_points"][0]), where
    assert np.array_equal(_sha(gts), fix["grid_timestamps_sha256"]), where + ": grid timestamps"
    assert np.array_equal(_sha(gval), fix["grid_values_sha256"]), where + ": grid values"
    count, mn, mx, sm = agg
    assert np.array_equal(count, fix["agg_count"]), where
    assert_f32_bits_equal(mn, fix["agg_min"], where + " min")
    assert_f32_bits_equal(mx, fix["agg_max"], where + " max")
    want = fix["agg_sum"]
    if sum_exact:
        assert np.array_equal(np.asarray(sm).view(np.uint64), want.view(np.uint64)), where
    else:  # the GPU folds a group's rows in a fixed in-order tree, the oracle left to right: 1e-12 relative
        both_nan = np.isnan(sm) & np.isnan(want)
        ok = both_nan | (np.abs(sm - want) <= 1e-12 * np.maximum(1.0, np.abs(want))) | (sm == want)
        assert ok.all(), (where, sm[~ok][:3], want[~ok][:3])


def test_fixtures_exist():
    assert len(FILES) >= 12


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_oracle_reproduces_golden_fixture(oracle, path):
    fix = np.load(path)
    ebs = list(zip(fix["in_eb_kind"].tolist(), fix["in_eb_value"].tolist()))
    seg = oracle.compress(fix["in_timestamps"], fix["in_values"], fix["in_unit_off"], eb=ebs)
    gts, gval, _ = oracle.grid(seg)
    _check(fix, seg, gts, gval, oracle.aggregate(seg, seg.unit_seg_off), os.path.basename(path), sum_exact=True)


@pytest.mark.gpu
@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_cuda_path_reproduces_golden_fixture(path):
    from modelardb_rs_b200 import compression as mc
    fix = np.load(path)
    ctx = mc.default_context()
    ebs = [mc.ErrorBound(int(k), float(v)) for k, v in zip(fix["in_eb_kind"], fix["in_eb_value"])]
    dev = mc.compress(fix["in_timestamps"], fix["in_values"], fix["in_unit_off"], ebs, ctx)
    seg = dev.to_host()
    gts, gval = mc.grid(seg, ctx=ctx)
    _check(fix, seg, gts, gval, mc.aggregate(seg, seg.unit_seg_off, ctx), os.path.basename(path), sum_exact=False)
    dev.free()
