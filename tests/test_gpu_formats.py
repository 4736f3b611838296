"""Parquet file -> segments -> CUDA grid / aggregates -> Arrow IPC stream, against the oracle:
what a query over a time series table does on either side of the hot path."""
import numpy as np
import pyarrow as pa
import pytest

from modelardb_rs_b200 import compression as mc
from modelardb_rs_b200 import formats as mf
from modelardb_rs_b200 import operators as ops
from modelardb_rs_b200 import synthetic as syn
from tests.parity_cases import assert_f32_bits_equal

pytestmark = pytest.mark.gpu


def _series(n_series=4, n=30000, seed=21, irregular=False):
    return syn.multi_series(n_series, n, seed, "sine", irregular=irregular)


def test_compress_to_parquet_equals_the_oracles_segments(oracle, tmp_path):
    ts, vals, off = _series()
    seg = mc.compress(ts, vals, off, mc.ErrorBound.try_new_relative(1.0))
    want = oracle.compress(ts, vals, off, eb=(2, 1.0))
    uso = seg.to_host().unit_seg_off
    tags = np.repeat(np.asarray(["s%d" % u for u in range(4)], object), np.diff(uso).astype(np.int64))
    path = str(tmp_path / "table.parquet")
    mf.write_segments(path, seg, field_column=1, tags={"series": tags})
    parts = list(mf.read_segments(path))
    lo = 0
    for back, extra in parts:
        for c in mc._COLUMNS:
            a, b = getattr(back, c), getattr(want, c)
            if c.endswith("_off"):
                assert np.array_equal(a, b[lo:lo + len(back) + 1] - b[lo]), c
            elif c.endswith("_data"):
                o = getattr(want, c[:-5] + "_off")
                assert a.tobytes() == b[int(o[lo]):int(o[lo + len(back)])].tobytes(), c
            else:
                assert a.tobytes() == b[lo:lo + len(back)].tobytes(), c
        assert list(extra["series"]) == list(tags[lo:lo + len(back)])
        lo += len(back)
    assert lo == len(want)


@pytest.mark.parametrize("irregular", [False, True])
def test_parquet_to_grid_to_ipc_stream(oracle, tmp_path, irregular):
    ts, vals, off = _series(irregular=irregular)
    want = oracle.compress(ts, vals, off, eb=(1, 0.25))
    want_ts, want_val, point_off = oracle.grid(want)
    tags = np.repeat(np.asarray(["s%d" % u for u in range(4)], object), np.diff(want.unit_seg_off).astype(np.int64))
    path = str(tmp_path / "table.parquet")
    mf.write_segments(path, mc.HostSegments(**{c: getattr(want, c) for c in mc._COLUMNS}), tags={"series": tags})
    stream = ops.GridStream(((s, [extra["series"]]) for s, extra in mf.read_segments(path)), 8192, n_tag_columns=1)
    wire = mf.send_query_result(stream, ["series"])
    got = mf.read_query_result(wire)
    assert all(b.schema.equals(mf.grid_schema(["series"])) and b.num_rows <= 8192 for b in got)
    got_ts = np.concatenate([b.column("timestamp").cast(pa.int64()).to_numpy() for b in got])
    got_val = np.concatenate([b.column("value").to_numpy() for b in got])
    assert np.array_equal(got_ts, want_ts) and np.array_equal(got_ts, ts)
    assert_f32_bits_equal(got_val, want_val, "values on the wire")
    want_tag = np.repeat(tags, np.diff(point_off).astype(np.int64))
    assert [t for b in got for t in b.column("series").to_pylist()] == list(want_tag)


def test_tag_runs_and_grouped_aggregates_on_the_device(oracle):
    ts, vals, off = _series(n_series=5, n=20000, seed=33)
    seg = mc.compress(ts, vals, off, mc.ErrorBound.try_new_relative(5.0))
    host = seg.to_host(copy=True)
    tags = np.repeat(np.asarray(["s%d" % u for u in range(5)], object), np.diff(host.unit_seg_off).astype(np.int64))
    out = list(ops.GridStream([(host, [tags])], 7000, n_tag_columns=1, tag_runs=True))
    assert np.array_equal(np.concatenate([b[0] for b in out]), ts)
    expanded = np.concatenate([np.repeat(v, l) for _, _, (v, l) in out])
    assert np.array_equal(expanded, np.repeat(np.asarray(["s%d" % u for u in range(5)], object), 20000))
    keys, count, mn, mx, sm = ops.grouped_model_aggregates(host, [tags])
    want = oracle.compress(ts, vals, off, eb=(2, 5.0))
    c, a, b, s = oracle.aggregate(want, want.unit_seg_off)
    assert keys == [("s%d" % u,) for u in range(5)]
    assert np.array_equal(count, c) and mn.tobytes() == a.tobytes() and mx.tobytes() == b.tobytes()
    assert np.all(np.abs(sm - s) <= 1e-12 * np.abs(s))
