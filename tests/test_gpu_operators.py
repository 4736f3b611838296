"""The host-side operator mirrors (modelardb_rs_b200/operators.py) against what the reference's operators compute:
GridStream = concatenation of grid() over all rows, cut into batch_size pieces (grid_exec.rs:261-430); the model
accumulators = folds of len() / sum() / min / max over all rows (model_simple_aggregates.rs:336-618)."""
import numpy as np
import pytest

from modelardb_rs_b200 import compression as mc
from modelardb_rs_b200 import operators as ops
from modelardb_rs_b200 import synthetic as syn
from tests.parity_cases import assert_f32_bits_equal

pytestmark = pytest.mark.gpu


def _segment_batches(oracle, n_batches=3):
    """A few segment batches (one per 'input RecordBatch') with one tag column, plus the oracle's points for them."""
    batches, want_ts, want_val, want_tag, segs = [], [], [], [], []
    for b in range(n_batches):
        ts, vals, off = syn.multi_series(3, 2500 + 700 * b, 40 + b, "sine" if b % 2 == 0 else "walk", irregular=(b == 2))
        eb = [(2, 1.0), (0, 0.0), (1, 0.5)][b]
        seg = oracle.compress(ts, vals, off, eb=eb)
        tags = np.array([f"series-{b}-{u}" for u in range(3) for _ in range(int(seg.unit_seg_off[u + 1] - seg.unit_seg_off[u]))], object)
        gts, gval, point_off = oracle.grid(seg)
        host = mc.HostSegments(**{c: getattr(seg, c) for c in mc._COLUMNS})
        batches.append((host, [tags]))
        segs.append(seg)
        want_ts.append(gts)
        want_val.append(gval)
        want_tag.append(np.repeat(tags, np.diff(point_off).astype(np.int64)))
    return batches, np.concatenate(want_ts), np.concatenate(want_val), np.concatenate(want_tag), segs


@pytest.mark.parametrize("batch_size", [1000, 8192, 1_000_000])
def test_grid_stream_equals_row_wise_grid(oracle, batch_size):
    batches, want_ts, want_val, want_tag, _ = _segment_batches(oracle)
    stream = ops.GridStream(batches, batch_size, n_tag_columns=1)
    out = list(stream)
    assert all(len(b[0]) <= batch_size for b in out)
    got_ts = np.concatenate([b[0] for b in out])
    got_val = np.concatenate([b[1] for b in out])
    got_tag = np.concatenate([b[2] for b in out])
    assert np.array_equal(got_ts, want_ts)
    assert_f32_bits_equal(got_val, want_val, "grid stream values")
    assert np.array_equal(got_tag, want_tag)
    assert stream.metrics.rows_created == len(want_ts)
    assert sum(stream.metrics.rows_by_model_type.values()) == len(want_ts)


def test_grid_stream_prunes_by_time_after_reconstruction(oracle):
    batches, want_ts, want_val, _, _ = _segment_batches(oracle)
    lo, hi = np.quantile(want_ts, [0.3, 0.6]).astype(np.int64)
    stream = ops.GridStream(batches, 4096, n_tag_columns=1, predicate=lambda t, v: (t >= lo) & (t <= hi))
    out = list(stream)
    keep = (want_ts >= lo) & (want_ts <= hi)
    assert np.array_equal(np.concatenate([b[0] for b in out]), want_ts[keep])
    assert_f32_bits_equal(np.concatenate([b[1] for b in out]), want_val[keep], "filtered values")


def test_grid_stream_skips_segments_outside_the_time_range(oracle):
    """Push-down of the time predicate to whole segments: same output as pruning after reconstruction,
    but segments that cannot contain a matching point are never reconstructed."""
    batches, want_ts, want_val, want_tag, _ = _segment_batches(oracle)
    lo, hi = np.quantile(want_ts, [0.45, 0.55]).astype(np.int64)
    pred = lambda t, v: (t >= lo) & (t <= hi)  # noqa: E731
    pruned = ops.GridStream(batches, 4096, n_tag_columns=1, predicate=pred)
    pushed = ops.GridStream(batches, 4096, n_tag_columns=1, predicate=pred, time_range=(int(lo), int(hi)))
    a, b = list(pruned), list(pushed)
    for col in range(3):
        assert np.array_equal(np.concatenate([x[col] for x in a]), np.concatenate([x[col] for x in b]))
    assert pushed.segments_skipped > 0
    assert pushed.metrics.rows_created < pruned.metrics.rows_created
    # a range that matches nothing: every segment is skipped and the stream is empty
    nothing = ops.GridStream(batches, 4096, n_tag_columns=1, predicate=lambda t, v: t < 0, time_range=(None, -1))
    out = list(nothing)  # (like the reference, a poll whose input batch yields no rows returns an empty batch)
    assert all(len(x[0]) == 0 for x in out) and nothing.segments_skipped == sum(len(h) for h, _ in batches)


def test_model_accumulators_equal_folds_over_rows(oracle):
    batches, want_ts, _, _, segs = _segment_batches(oracle)
    accs = [ops.ModelCountAccumulator(), ops.ModelMinAccumulator(), ops.ModelMaxAccumulator(), ops.ModelSumAccumulator(), ops.ModelAvgAccumulator()]
    for host, _ in batches:
        for a in accs:
            a.update_batch(host)
    count, mn, mx, sm, avg = [a.state() for a in accs]
    # the reference folds row by row: len() into i64, min/max NaN-ignoring from f32::MAX / f32::MIN, f32 row sums into f64
    want_count, want_sum = 0, 0.0
    want_min, want_max = ops.F32_MAX, ops.F32_MIN
    for seg in segs:
        sums = oracle.segment_sums(seg)
        want_count += int(oracle.grid_count(seg)[-1])
        for s in sums:
            want_sum += float(s)
        for v in seg.min_value:
            want_min = v if v < want_min else want_min
        for v in seg.max_value:
            want_max = v if v > want_max else want_max
    assert count == [want_count] == [len(want_ts)]
    assert mn[0] == want_min and mx[0] == want_max
    assert abs(sm[0] - want_sum) <= 1e-12 * abs(want_sum)  # in-order tree per batch vs left fold
    assert avg[0] == want_count and abs(avg[1] - want_sum) <= 1e-12 * abs(want_sum)
    # state() resets (model_simple_aggregates.rs:372-376)
    assert accs[0].state() == [0] and accs[1].state() == [ops.F32_MAX] and accs[3].state() == [0.0]
    with pytest.raises(RuntimeError):
        accs[0].evaluate()
