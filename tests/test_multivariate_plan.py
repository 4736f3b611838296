"""The host part of try_compress_multivariate_time_series (modelardb_rs_b200/compression.py: plan_multivariate) against
a literal restatement of the reference's loop (compression.rs:42-141): sort by tags then time, cut where a tag changes,
one unit per (series, field) in series-major order.  CPU only; the GPU test compresses the planned units."""
import numpy as np
import pytest

from modelardb_rs_b200 import compression as mc


def _reference_loop(ts, tags, fields):
    rows = sorted(range(len(ts)), key=lambda i: tuple(t[i] for t in tags) + (ts[i],))
    units, cur, cur_tags = [], [], None
    for i in rows:
        key = tuple(t[i] for t in tags)
        if cur_tags is not None and key != cur_tags:
            units += [(cur_tags, f, [ts[j] for j in cur], [fields[f][j] for j in cur]) for f in range(len(fields))]
            cur = []
        cur_tags = key
        cur.append(i)
    if cur:
        units += [(cur_tags, f, [ts[j] for j in cur], [fields[f][j] for j in cur]) for f in range(len(fields))]
    return units


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_plan_matches_the_reference_loop(seed):
    rng = np.random.default_rng(seed)
    n = 400
    tag_a = rng.choice(["north", "south", "east", "Zeta", "älv"], n)
    tag_b = rng.choice(["t1", "t10", "t2"], n)
    ts = rng.permutation(n).astype(np.int64) * 7 + 1000  # distinct timestamps, shuffled rows
    fields = [rng.standard_normal(n).astype(np.float32), rng.uniform(0, 1, n).astype(np.float32)]
    u_ts, u_val, off, u_series, u_field, series_tags = mc.plan_multivariate(ts, [tag_a, tag_b], fields)
    want = _reference_loop(ts.tolist(), [tag_a.tolist(), tag_b.tolist()], [f.tolist() for f in fields])
    assert len(want) == len(u_series)
    for u, (tags, f, wts, wval) in enumerate(want):
        a, b = int(off[u]), int(off[u + 1])
        assert series_tags[u_series[u]] == tags and u_field[u] == f
        assert u_ts[a:b].tolist() == wts
        assert u_val[a:b].tolist() == pytest.approx(wval, abs=0)


def test_plan_edge_cases():
    empty = mc.plan_multivariate(np.zeros(0, np.int64), [np.zeros(0, str)], [np.zeros(0, np.float32)])
    assert len(empty[3]) == 0 and empty[2].tolist() == [0]
    one = mc.plan_multivariate(np.array([5, 3, 4]), [], [np.array([1.0, 2.0, 3.0], np.float32)])  # no tags: one series
    assert one[0].tolist() == [3, 4, 5] and one[1].tolist() == [2.0, 3.0, 1.0] and one[5] == [()]
    with pytest.raises(ValueError):
        mc.plan_multivariate(np.arange(3), [["a", "b"]], [np.zeros(3, np.float32)])


@pytest.mark.gpu
def test_multivariate_compress_matches_per_series_oracle(oracle):
    rng = np.random.default_rng(5)
    n_series, n = 5, 3000
    ts = np.tile(1_600_000_000_000_000 + 1000 * np.arange(n, dtype=np.int64), n_series)
    tag = np.repeat([f"turbine-{k}" for k in range(n_series)], n)
    f0 = (100 + 10 * np.sin(np.arange(n_series * n) / 300.0) + 0.1 * rng.standard_normal(n_series * n)).astype(np.float32)
    f1 = rng.uniform(-1, 1, n_series * n).astype(np.float32)
    shuffle = rng.permutation(n_series * n)
    ebs = [mc.ErrorBound(2, 1.0), mc.ErrorBound(0, 0.0)]
    got = mc.try_compress_multivariate_time_series(ts[shuffle], [tag[shuffle]], [f0[shuffle], f1[shuffle]], ebs)
    assert len(got) == n_series * 2
    from tests.parity_cases import assert_segments_equal
    for k, (tags, field, seg) in enumerate(got):
        s, f = divmod(k, 2)
        assert tags == (f"turbine-{s}",) and field == f
        sl = slice(s * n, (s + 1) * n)
        want = oracle.compress(ts[sl], (f0, f1)[f][sl], eb=[(2, 1.0), (0, 0.0)][f])
        assert_segments_equal(seg, want, f"series {s} field {f}")


@pytest.mark.gpu
@pytest.mark.parametrize("n,n_codes,ts_kind", [(1, 1, "distinct"), (257, 3, "distinct"), (5000, 70000, "dups"), (100_003, 300, "distinct"),
                                               (40_000, 1, "equal"), (33_333, 1000, "negative"), (2_000_000, 5000, "dups")])
def test_device_sort_equals_numpy_lexsort(n, n_codes, ts_kind):
    """mdbcu_sort_rows: the stable radix sort by (tag tuple code, timestamp) against np.lexsort (also stable), for sizes
    around the block boundaries, codes needing one to three bytes, duplicate / equal / negative timestamps; and
    mdbcu_take_rows against numpy's fancy indexing."""
    rng = np.random.default_rng(n + n_codes)
    code = rng.integers(0, n_codes, n).astype(np.uint32)
    if ts_kind == "distinct":
        ts = (rng.permutation(n).astype(np.int64) * 13 + 1_700_000_000_000)
    elif ts_kind == "dups":
        ts = rng.integers(0, max(2, n // 50), n).astype(np.int64) + 1_700_000_000_000_000
    elif ts_kind == "equal":
        ts = np.full(n, 42, np.int64)
    else:
        ts = rng.integers(-2 ** 62, 2 ** 62, n).astype(np.int64)
    ctx = mc.Context(0)
    order = mc.sort_rows(code, ts, ctx)
    want = np.lexsort([ts, code])
    assert np.array_equal(order, want.astype(np.uint32))
    f0, f1 = rng.standard_normal(n).astype(np.float32), rng.uniform(0, 1, n).astype(np.float32)
    t_out, (g0, g1) = mc.take_rows(order, ts, [f0, f1], ctx)
    assert np.array_equal(t_out, ts[want]) and np.array_equal(g0, f0[want]) and np.array_equal(g1, f1[want])
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_device_plan_equals_host_plan(seed):
    """plan_multivariate with the sort and the gather on the device gives the arrays of the host plan (which is checked
    against the reference's loop above)."""
    rng = np.random.default_rng(seed)
    n = 20_000
    tag_a = rng.choice(["north", "south", "east", "Zeta", "älv"], n)
    tag_b = rng.choice(["t1", "t10", "t2"], n)
    ts = rng.permutation(n).astype(np.int64) * 7 + 1000
    fields = [rng.standard_normal(n).astype(np.float32), rng.uniform(0, 1, n).astype(np.float32)]
    ctx = mc.Context(0)
    host = mc.plan_multivariate(ts, [tag_a, tag_b], fields)
    dev = mc.plan_multivariate(ts, [tag_a, tag_b], fields, device_ctx=ctx)
    for a, b in zip(host[:5], dev[:5]):
        assert np.array_equal(a, b)
    assert host[5] == dev[5]
    ctx.close()
