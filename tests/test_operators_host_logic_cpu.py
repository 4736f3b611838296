"""The HOST logic of modelardb_rs_b200/operators.py (leftovers, batch_size slicing, tag runs, predicate, time-range
push-down, grouping by tags) on a machine without a GPU: the three calls that would cross the C-ABI are replaced by the
oracle for the duration of a test.  The same operators run against the CUDA library in tests/test_gpu_operators.py."""
import numpy as np
import pytest

from modelardb_rs_b200 import compression as mc
from modelardb_rs_b200 import operators as ops
from modelardb_rs_b200 import synthetic as syn


@pytest.fixture
def oracle_behind_the_cabi(oracle, monkeypatch):
    def as_oracle(host):
        return oracle.Segments(**{c: getattr(host, c) for c in mc._COLUMNS})

    def grid_count(host, ctx=None):
        off = oracle.grid_count(as_oracle(host))
        return off, int(off[-1])

    def grid(host, ctx=None):
        ts, val, _ = oracle.grid(as_oracle(host))
        return ts, val

    def aggregate(host, group_off=None, ctx=None):
        return oracle.aggregate(as_oracle(host), group_off)

    monkeypatch.setattr(mc, "grid_count", grid_count)
    monkeypatch.setattr(mc, "grid", grid)
    monkeypatch.setattr(mc, "aggregate", aggregate)
    return oracle


def _batches(oracle, n_batches=3, n_series=3):
    batches, want_ts, want_val, want_tag, hosts = [], [], [], [], []
    for b in range(n_batches):
        ts, vals, off = syn.multi_series(n_series, 2500 + 700 * b, 40 + b, "sine" if b % 2 == 0 else "walk", irregular=(b == 2))
        seg = oracle.compress(ts, vals, off, eb=[(2, 1.0), (0, 0.0), (1, 0.5)][b % 3])
        tags = np.array([f"series-{u}" for u in range(n_series) for _ in range(int(seg.unit_seg_off[u + 1] - seg.unit_seg_off[u]))], object)
        gts, gval, point_off = oracle.grid(seg)
        host = mc.HostSegments(**{c: getattr(seg, c) for c in mc._COLUMNS})
        batches.append((host, [tags]))
        hosts.append(host)
        want_ts.append(gts)
        want_val.append(gval)
        want_tag.append(np.repeat(tags, np.diff(point_off).astype(np.int64)))
    return batches, np.concatenate(want_ts), np.concatenate(want_val), np.concatenate(want_tag), hosts


@pytest.mark.parametrize("batch_size", [1, 999, 8192, 10_000_000])
def test_stream_is_the_concatenated_grid(oracle_behind_the_cabi, batch_size):
    batches, want_ts, want_val, want_tag, _ = _batches(oracle_behind_the_cabi, n_batches=2 if batch_size == 1 else 3)
    if batch_size == 1:
        batches = [(h.slice(0, 3), [t[0][:3]]) for h, t in batches]
        want = [oracle_behind_the_cabi.grid(oracle_behind_the_cabi.Segments(**{c: getattr(h, c) for c in mc._COLUMNS})) for h, _ in batches]
        want_ts, want_val = np.concatenate([w[0] for w in want]), np.concatenate([w[1] for w in want])
        want_tag = np.concatenate([np.repeat(t[0], np.diff(w[2]).astype(np.int64)) for (_, t), w in zip(batches, want)])
    out = list(ops.GridStream(batches, batch_size, n_tag_columns=1))
    assert all(len(b[0]) <= batch_size and len(b[0]) == len(b[1]) == len(b[2]) for b in out)
    assert np.array_equal(np.concatenate([b[0] for b in out]), want_ts)
    assert np.concatenate([b[1] for b in out]).tobytes() == want_val.tobytes()
    assert np.array_equal(np.concatenate([b[2] for b in out]), want_tag)


def test_tag_runs_expand_to_the_repeated_column(oracle_behind_the_cabi):
    batches, want_ts, _, want_tag, _ = _batches(oracle_behind_the_cabi)
    for predicate in (None, lambda t, v: (t % 7000) < 3000):
        out = list(ops.GridStream(batches, 3000, n_tag_columns=1, tag_runs=True, predicate=predicate))
        expanded = []
        for ts, vals, (values, lens) in out:
            assert int(lens.sum()) == len(ts) and (lens > 0).all() and len(values) == len(lens)
            expanded.append(np.repeat(values, lens))
        keep = np.ones(len(want_ts), bool) if predicate is None else predicate(want_ts, None)
        assert np.array_equal(np.concatenate(expanded), want_tag[keep])
        assert np.array_equal(np.concatenate([b[0] for b in out]), want_ts[keep])
        # far fewer runs than rows: that is the point of the run-length form
        assert sum(len(b[2][0]) for b in out) < len(want_ts) // 4


def test_predicate_and_time_range_agree(oracle_behind_the_cabi):
    batches, want_ts, want_val, want_tag, hosts = _batches(oracle_behind_the_cabi)
    lo, hi = np.quantile(want_ts, [0.45, 0.55]).astype(np.int64)
    pred = lambda t, v: (t >= lo) & (t <= hi)  # noqa: E731
    pruned = ops.GridStream(batches, 4096, n_tag_columns=1, predicate=pred)
    pushed = ops.GridStream(batches, 4096, n_tag_columns=1, predicate=pred, time_range=(int(lo), int(hi)))
    a, b = list(pruned), list(pushed)
    keep = pred(want_ts, None)
    for col, want in enumerate((want_ts[keep], want_val[keep], want_tag[keep])):
        assert np.array_equal(np.concatenate([x[col] for x in a]), want)
        assert np.array_equal(np.concatenate([x[col] for x in b]), want)
    assert pushed.segments_skipped > 0 and pushed.metrics.rows_created < pruned.metrics.rows_created
    assert pruned.metrics.rows_created == len(want_ts)
    nothing = ops.GridStream(batches, 4096, n_tag_columns=1, predicate=lambda t, v: t < 0, time_range=(None, -1))
    assert all(len(x[0]) == 0 for x in nothing) and nothing.segments_skipped == sum(len(h) for h in hosts)


def test_no_tag_columns_and_mismatched_tags(oracle_behind_the_cabi):
    batches, want_ts, _, _, _ = _batches(oracle_behind_the_cabi, n_batches=1)
    out = list(ops.GridStream([(batches[0][0], [])], 5000))
    assert all(len(b) == 2 for b in out) and np.array_equal(np.concatenate([b[0] for b in out]), want_ts)
    with pytest.raises(ValueError, match="same tag columns"):
        list(ops.GridStream(batches, 5000, n_tag_columns=2))
    with pytest.raises(ValueError, match="batch_size"):
        ops.GridStream(batches, 0)


def test_grouped_aggregates_equal_aggregates_of_the_points(oracle_behind_the_cabi):
    """GROUP BY tag over segments = the aggregates DataFusion computes over the reconstructed points of each group:
    COUNT / MIN / MAX equal, SUM within the 0.001 % the reference's own test allows (integration_test.rs:1128-1171)."""
    oracle = oracle_behind_the_cabi
    batches, _, _, _, hosts = _batches(oracle, n_batches=2, n_series=4)
    for host, (tags,) in batches:
        second = np.array([t[-1] for t in tags], object)  # a second tag column, constant within a series
        keys, count, mn, mx, sm = ops.grouped_model_aggregates(host, [tags, second])
        assert keys == [(f"series-{u}", str(u)) for u in range(4)]
        ts, val, point_off = oracle.grid(oracle.Segments(**{c: getattr(host, c) for c in mc._COLUMNS}))
        point_tag = np.repeat(tags, np.diff(point_off).astype(np.int64))
        for k, key in enumerate(keys):
            points = val[point_tag == key[0]]
            assert count[k] == len(points)
            assert mn[k] == points.min() and mx[k] == points.max()
            want = float(points.astype(np.float64).sum())
            assert abs(sm[k] - want) <= 1e-5 * abs(want)


def test_grouped_aggregates_merge_repeated_keys(oracle_behind_the_cabi):
    """Rows of one series that are not adjacent (two files of the same series) fold into one group, in row order."""
    oracle = oracle_behind_the_cabi
    batches, _, _, _, _ = _batches(oracle, n_batches=1, n_series=3)
    host, (tags,) = batches[0]
    order = np.concatenate([np.flatnonzero(tags == "series-1"), np.flatnonzero(tags == "series-0"), np.flatnonzero(tags == "series-2"),
                            np.flatnonzero(tags == "series-1")[:5]])
    shuffled, shuffled_tags = _take_any(host, order), tags[order]
    keys, count, mn, mx, sm = ops.grouped_model_aggregates(shuffled, [shuffled_tags])
    assert keys == [("series-1",), ("series-0",), ("series-2",)]
    for k, (name,) in enumerate(keys):
        rows = np.flatnonzero(shuffled_tags == name)
        c, a, b, s = oracle.aggregate(oracle.Segments(**{col: getattr(_take_any(shuffled, rows), col) for col in mc._COLUMNS}))
        assert count[k] == c[0] and mn[k] == a[0] and mx[k] == b[0] and abs(sm[k] - s[0]) <= 1e-12 * abs(s[0])
    assert ops.grouped_model_aggregates(host.slice(0, 0), [tags[:0]])[0] == []
    with pytest.raises(ValueError, match="one value per segment"):
        ops.grouped_model_aggregates(host, [tags[:-1]])


def _take_any(host, order):
    """Rows in ANY order (HostSegments.take keeps the original order): one slice per row, concatenated."""
    parts = [host.slice(int(i), int(i) + 1) for i in order]
    cols = {c: np.concatenate([getattr(p, c) for p in parts]) for c in ("model_type_id", "start_time", "end_time", "min_value", "max_value")}
    for name in ("timestamps", "values", "residuals"):
        lens = np.array([len(getattr(p, name + "_data")) for p in parts], np.uint64)
        cols[name + "_off"] = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
        cols[name + "_data"] = np.concatenate([getattr(p, name + "_data") for p in parts]) if len(parts) else np.zeros(0, np.uint8)
    return mc.HostSegments(**cols)


@pytest.mark.parametrize("limit", [1, 7, 4096, 5000, 20000])
def test_limit_gives_the_first_rows_and_skips_the_rest(oracle_behind_the_cabi, limit):
    """LIMIT: batch_size = min(limit, batch_size) as in the reference (grid_exec.rs:239-246); the stream also ends after
    `limit` rows and only reconstructs the segments it needs for them."""
    batches, want_ts, want_val, want_tag, hosts = _batches(oracle_behind_the_cabi)
    stream = ops.GridStream(batches, 4096, n_tag_columns=1, limit=limit)
    out = list(stream)
    assert stream.batch_size == min(limit, 4096) and all(len(b[0]) <= stream.batch_size for b in out)
    n = min(limit, len(want_ts))
    assert np.array_equal(np.concatenate([b[0] for b in out]), want_ts[:n])
    assert np.concatenate([b[1] for b in out]).tobytes() == want_val[:n].tobytes()
    assert np.array_equal(np.concatenate([b[2] for b in out]), want_tag[:n])
    if limit < len(want_ts) // 2:
        assert stream.segments_skipped > 0 and stream.metrics.rows_created < len(want_ts)
    # with a predicate nothing can be skipped up front, but the first `limit` surviving rows are the same
    keep = (want_ts % 5000) < 2500
    filtered = list(ops.GridStream(batches, 4096, n_tag_columns=1, limit=limit, predicate=lambda t, v: (t % 5000) < 2500))
    m = min(limit, int(keep.sum()))
    assert np.array_equal(np.concatenate([b[0] for b in filtered]), want_ts[keep][:m])
    with pytest.raises(ValueError, match="limit"):
        ops.GridStream(batches, 4096, limit=0)
