"""modelardb_rs_b200/formats.py: the reference's schemas, Parquet writer settings and IPC framing around the segment
columns of the C-ABI.  CPU only: segments come from the oracle."""
import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq
import pytest

from modelardb_rs_b200 import compression as mc
from modelardb_rs_b200 import formats as mf
from modelardb_rs_b200 import synthetic as syn


def _host(oracle, kind="walk", eb=(2, 0.5), n_series=3, n=4000):
    ts, vals, off = syn.multi_series(n_series, n, 8, kind)
    seg = oracle.compress(ts, vals, off, eb=eb)
    return mc.HostSegments(**{c: getattr(seg, c) for c in mc._COLUMNS})


def _same(a: mc.HostSegments, b: mc.HostSegments):
    for c in mc._COLUMNS:
        x, y = getattr(a, c), getattr(b, c)
        assert x.dtype == y.dtype and x.shape == y.shape, c
        assert x.tobytes() == y.tobytes(), c  # bit patterns, NaN payloads included


def test_schemas_are_the_references():
    q = mf.query_compressed_schema()
    assert q.names == ["model_type_id", "start_time", "end_time", "timestamps", "min_value", "max_value", "values", "residuals", "error"]
    assert [str(t) for t in q.types] == ["int8", "timestamp[us]", "timestamp[us]", "binary_view", "float", "float", "binary_view",
                                         "binary_view", "float"]
    assert not any(f.nullable for f in q)
    # COMPRESSED_METADATA_SIZE_IN_BYTES (schemas.rs:57-64): the fixed-width columns add up to 29 bytes
    assert sum(t.bit_width // 8 for t in q.types if pa.types.is_primitive(t)) == 29
    c = mf.compressed_schema(["location", "install"])
    assert c.names[9:] == ["field_column", "location", "install"]
    assert str(c.field("field_column").type) == "int16" and str(c.field("location").type) == "string_view"
    assert mf.grid_schema(["tag"]).names == ["timestamp", "value", "tag"]


@pytest.mark.parametrize("kind,eb", [("walk", (2, 0.5)), ("sine", (2, 1.0)), ("mixed", (0, 0.0)), ("sine", (1, 0.05))])
def test_record_batch_round_trip(oracle, kind, eb):
    host = _host(oracle, kind, eb)
    batch = mf.segments_to_record_batch(host, field_column=3, tags={"location": "aalborg"})
    assert batch.schema.equals(mf.compressed_schema(["location"]))
    assert batch.num_rows == len(host)
    assert np.isnan(batch.column("error").to_numpy()).all()
    # a row is the bytes the oracle produced
    i = len(host) // 2
    assert batch.column("values")[i].as_py() == host.row(i)["values"]
    assert batch.column("timestamps")[i].as_py() == host.row(i)["timestamps"]
    back, extra = mf.record_batch_to_segments(batch)
    _same(host, back)
    assert (extra["field_column"] == 3).all() and (extra["location"] == "aalborg").all()


def test_query_schema_without_tags(oracle):
    host = _host(oracle)
    batch = mf.segments_to_record_batch(host)
    assert batch.schema.equals(mf.query_compressed_schema())
    back, extra = mf.record_batch_to_segments(batch)
    _same(host, back)
    assert extra == {}


def test_empty_batch():
    empty = mc.HostSegments(model_type_id=[], start_time=[], end_time=[], min_value=[], max_value=[], timestamps_off=[0],
                            timestamps_data=[], values_off=[0], values_data=[], residuals_off=[0], residuals_data=[])
    batch = mf.segments_to_record_batch(empty)
    assert batch.num_rows == 0
    back, _ = mf.record_batch_to_segments(batch)
    _same(empty, back)


def test_sliced_batches_convert(oracle):
    """Arrow slices carry offsets into shared buffers; the conversion must rebase them."""
    host = _host(oracle)
    batch = mf.segments_to_record_batch(host)
    lo, hi = 7, len(host) - 3
    back, _ = mf.record_batch_to_segments(batch.slice(lo, hi - lo))
    _same(host.slice(lo, hi), back)


def test_accepts_plain_binary_columns(oracle):
    """Files written by other tools hold Binary / LargeBinary instead of BinaryView."""
    host = _host(oracle)
    batch = mf.segments_to_record_batch(host)
    table = pa.Table.from_batches([batch])
    for name in ("timestamps", "values", "residuals"):
        table = table.set_column(table.schema.get_field_index(name), name, table.column(name).cast(pa.binary()))
    back, _ = mf.record_batch_to_segments(table)
    _same(host, back)


def test_missing_column_is_an_error(oracle):
    batch = mf.segments_to_record_batch(_host(oracle))
    with pytest.raises(ValueError, match="missing: values"):
        mf.record_batch_to_segments(batch.drop_columns(["values"]))


def test_parquet_round_trip_and_writer_properties(oracle, tmp_path):
    host = _host(oracle, "sine", (2, 1.0), n_series=4, n=60000)
    path = str(tmp_path / "segments.parquet")
    per_row_tag = np.asarray(["series_%d" % (i % 4) for i in range(len(host))], object)
    mf.write_segments(path, host, field_column=1, tags={"series": per_row_tag})
    # lib.rs:248-261: ZSTD, PLAIN, no dictionary, no statistics, <= 65 536 rows per row group, sorting column recorded
    meta = pq.ParquetFile(path).metadata
    assert meta.num_rows == len(host)
    for g in range(meta.num_row_groups):
        group = meta.row_group(g)
        assert group.num_rows <= 65536
        assert group.sorting_columns[0].column_index == 1 and not group.sorting_columns[0].descending
        for c in range(group.num_columns):
            column = group.column(c)
            assert column.compression == "ZSTD"
            assert not column.has_dictionary_page
            assert "PLAIN" in column.encodings and not any("DICTIONARY" in e for e in column.encodings)
            assert not column.is_stats_set
    parts = list(mf.read_segments(path))
    assert len(parts) == meta.num_row_groups
    lo = 0
    for back, extra in parts:
        _same(host.slice(lo, lo + len(back)), back)
        assert list(extra["series"]) == list(per_row_tag[lo:lo + len(back)])
        assert (extra["field_column"] == 1).all()
        lo += len(back)
    assert lo == len(host)


def test_parquet_many_row_groups(oracle, tmp_path):
    """More than 65 536 rows are cut into row groups; reading yields them in order."""
    ts, vals, off = syn.multi_series(1, 300000, 3, "sine")
    seg = oracle.compress(ts, vals, off, eb=(2, 0.05))  # a tight bound: many short rows
    host = mc.HostSegments(**{c: getattr(seg, c) for c in mc._COLUMNS})
    reps = 65536 // len(host) + 2
    rows = np.tile(np.arange(len(host)), reps)
    big = host.take(rows) if len(host) * reps > 65536 else host
    assert len(big) > 65536
    path = str(tmp_path / "big.parquet")
    mf.write_segments(path, big)
    parts = [p for p, _ in mf.read_segments(path)]
    assert len(parts) >= 2 and all(len(p) <= 65536 for p in parts)
    lo = 0
    for p in parts:
        _same(big.slice(lo, lo + len(p)), p)
        lo += len(p)


def test_parquet_extension_and_bad_file(tmp_path, oracle):
    batch = mf.segments_to_record_batch(_host(oracle))
    with pytest.raises(ValueError, match="extension"):
        mf.write_record_batch_to_apache_parquet_file(str(tmp_path / "segments.txt"), batch)
    bad = tmp_path / "not.parquet"
    bad.write_bytes(b"this is not a parquet file")
    with pytest.raises(ValueError, match="not an Apache Parquet file"):
        list(mf.read_segments(str(bad)))
    with pytest.raises(ValueError, match="not an Apache Parquet file"):
        list(mf.read_segments(str(tmp_path / "absent.parquet")))


def test_query_result_stream(oracle):
    """remote.rs:169-211: schema message first, then one message per batch; values keep their bit patterns."""
    ts, vals, off = syn.multi_series(2, 5000, 5, "sine")
    seg = oracle.compress(ts, vals, off, eb=(2, 1.0))
    g_ts, g_val, _ = oracle.grid(seg)
    tags = np.repeat(np.asarray(["a", "b"], object), 5000)
    batches = [(g_ts[lo:lo + 4096], g_val[lo:lo + 4096], tags[lo:lo + 4096]) for lo in range(0, len(g_ts), 4096)]
    stream = mf.send_query_result(batches, ["tag"])
    reader = pa.ipc.open_stream(stream)
    assert reader.schema.equals(mf.grid_schema(["tag"]))
    got = mf.read_query_result(stream)
    assert [b.num_rows for b in got] == [len(b[0]) for b in batches]
    all_ts = np.concatenate([b.column("timestamp").cast(pa.int64()).to_numpy() for b in got])
    all_val = np.concatenate([b.column("value").to_numpy() for b in got])
    assert np.array_equal(all_ts, g_ts) and all_val.tobytes() == g_val.tobytes()
    assert [t for b in got for t in b.column("tag").to_pylist()] == list(tags)
    # an empty result is a stream with only the schema
    assert mf.read_query_result(mf.send_query_result([], ["tag"])) == []
