"""HostSegments.slice / take (modelardb_rs_b200/compression.py): row selection with rebased binary columns.  CPU only."""
import numpy as np

from modelardb_rs_b200 import compression as mc
from modelardb_rs_b200 import synthetic as syn


def _host(oracle):
    ts, vals, off = syn.multi_series(3, 4000, 8, "walk")
    seg = oracle.compress(ts, vals, off, eb=(2, 0.5))
    return mc.HostSegments(**{c: getattr(seg, c) for c in mc._COLUMNS}), seg


def test_take_keeps_rows_bit_for_bit(oracle):
    host, seg = _host(oracle)
    rng = np.random.default_rng(1)
    mask = rng.random(len(host)) < 0.4
    sub = host.take(mask)
    idx = np.flatnonzero(mask)
    assert len(sub) == len(idx)
    for k, i in enumerate(idx):
        a, b = sub.row(k), host.row(int(i))
        assert a.keys() == b.keys()
        for key in a:
            if isinstance(a[key], (bytes, int)):
                assert a[key] == b[key], key
            else:
                assert np.float32(a[key]).view(np.uint32) == np.float32(b[key]).view(np.uint32), key
    # indices instead of a mask, the empty selection, and agreement with slice on a contiguous range
    assert host.take(idx).row(0) == sub.row(0)
    assert len(host.take(np.zeros(len(host), bool))) == 0
    lo, hi = 5, min(40, len(host))
    a, b = host.take(np.arange(lo, hi)), host.slice(lo, hi)
    for c in mc._COLUMNS:
        assert np.array_equal(getattr(a, c), getattr(b, c)), c
