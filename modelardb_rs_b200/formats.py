"""The data formats on either side of the hot path, host side only:

  * segments <-> Arrow record batches with the reference's schemas
        QUERY_COMPRESSED_SCHEMA / COMPRESSED_SCHEMA      crates/modelardb_types/src/schemas.rs:31-52
        + one Utf8View column per tag                    crates/modelardb_types/src/types.rs (compressed_schema)
  * segments <-> Apache Parquet files written with the reference's writer properties
        apache_parquet_writer_properties                 crates/modelardb_storage/src/lib.rs:248-261
        read_record_batch_from_apache_parquet_file       crates/modelardb_storage/src/lib.rs:173-210
  * reconstructed data points -> Arrow IPC stream (schema message, then one message per record batch), what
        send_query_result                                crates/modelardb_server/src/remote.rs:169-211
    puts on the wire for a query result.

The file and wire encodings themselves are pyarrow's (Arrow C++): this module only fixes the schemas, the writer
settings and the zero-copy conversions between Arrow's binary columns and the offsets + data arrays of the C-ABI
(`mdbcu_segments_view`).  Nothing here computes on data points; the kernels stay behind modelardb_rs_b200.compression.
"""
from __future__ import annotations

import io
from typing import Dict, Iterable, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq

from . import compression as mc

FIELD_COLUMN = "field_column"  # schemas.rs:28
TIMESTAMP = pa.timestamp("us")  # ArrowTimestamp: TimestampMicrosecondType without a time zone
VALUE = pa.float32()  # ArrowValue

_BINARY_COLUMNS = ("timestamps", "values", "residuals")


def query_compressed_schema() -> pa.Schema:
    """QUERY_COMPRESSED_SCHEMA (schemas.rs:40-52); every field is non-nullable."""
    return pa.schema([
        pa.field("model_type_id", pa.int8(), False),
        pa.field("start_time", TIMESTAMP, False),
        pa.field("end_time", TIMESTAMP, False),
        pa.field("timestamps", pa.binary_view(), False),
        pa.field("min_value", VALUE, False),
        pa.field("max_value", VALUE, False),
        pa.field("values", pa.binary_view(), False),
        pa.field("residuals", pa.binary_view(), False),
        pa.field("error", pa.float32(), False),
    ])


def compressed_schema(tag_names: Sequence[str] = ()) -> pa.Schema:
    """COMPRESSED_SCHEMA (schemas.rs:31-36) followed by the table's tag columns as Utf8View."""
    fields = list(query_compressed_schema()) + [pa.field(FIELD_COLUMN, pa.int16(), False)]
    fields += [pa.field(name, pa.string_view(), False) for name in tag_names]
    return pa.schema(fields)


def grid_schema(tag_names: Sequence[str] = ()) -> pa.Schema:
    """GRID_SCHEMA (schemas.rs:67-72) followed by the tag columns GridExec appends (grid_exec.rs:126-143)."""
    fields = [pa.field("timestamp", TIMESTAMP, False), pa.field("value", VALUE, False)]
    return pa.schema(fields + [pa.field(name, pa.string_view(), False) for name in tag_names])


# ---- segments <-> record batches -------------------------------------------------------------------------------------

def _binary_array(off: np.ndarray, data: np.ndarray) -> pa.Array:
    """Rows data[off[i]:off[i+1]] as a BinaryView array.  The LargeBinary array over the two numpy buffers is zero-copy;
    the cast builds the 16-byte views the reference's schema asks for."""
    off = np.ascontiguousarray(off, dtype=np.int64)
    data = np.ascontiguousarray(data, dtype=np.uint8)
    large = pa.Array.from_buffers(pa.large_binary(), len(off) - 1, [None, pa.py_buffer(off), pa.py_buffer(data)])
    return large.cast(pa.binary_view())


def _offsets_and_data(column) -> Tuple[np.ndarray, np.ndarray]:
    """The inverse: any Arrow binary column (Binary, LargeBinary, BinaryView, possibly chunked) as offsets + data."""
    if isinstance(column, pa.ChunkedArray):
        column = column.combine_chunks() if column.num_chunks != 1 else column.chunk(0)
    if column.null_count:
        raise ValueError("binary columns of compressed segments are not nullable")
    large = column.cast(pa.large_binary())
    n = len(large)
    _, off_buf, data_buf = large.buffers()
    off = np.frombuffer(off_buf, dtype=np.int64, count=n + 1, offset=large.offset * 8) if n or off_buf is not None else np.zeros(1, np.int64)
    data = np.frombuffer(data_buf, dtype=np.uint8) if data_buf is not None else np.zeros(0, np.uint8)
    lo, hi = int(off[0]), int(off[-1])
    return (off - lo).astype(np.uint64), data[lo:hi]


def segments_to_record_batch(segments, field_column: Optional[int] = None,
                             tags: Optional[Dict[str, object]] = None) -> pa.RecordBatch:
    """A batch of segments as the record batch CompressedSegmentBatchBuilder::finish produces (types.rs:492-516): the
    nine query columns, then `field_column` (constant, if given) and one Utf8View column per tag.  A tag is either one
    string for the whole batch (what compress produces: one series per call) or one string per row."""
    host = segments.to_host() if isinstance(segments, mc.CompressedSegments) else segments
    n = len(host)
    arrays = [
        pa.array(host.model_type_id, pa.int8()),
        pa.array(host.start_time, pa.int64()).cast(TIMESTAMP),
        pa.array(host.end_time, pa.int64()).cast(TIMESTAMP),
        _binary_array(host.timestamps_off, host.timestamps_data),
        pa.array(host.min_value, VALUE),
        pa.array(host.max_value, VALUE),
        _binary_array(host.values_off, host.values_data),
        _binary_array(host.residuals_off, host.residuals_data),
        pa.array(np.full(n, np.nan, np.float32), pa.float32()),  # the error column is always NaN (compression.rs:398)
    ]
    schema = query_compressed_schema()
    if field_column is not None or tags:
        tags = tags or {}
        schema = compressed_schema(list(tags))
        arrays.append(pa.array(np.full(n, 0 if field_column is None else field_column, np.int16), pa.int16()))
        for value in tags.values():
            column = [value] * n if isinstance(value, str) else list(value)
            if len(column) != n:
                raise ValueError("a tag column needs one value per segment")
            arrays.append(pa.array(column, pa.string()).cast(pa.string_view()))
    return pa.RecordBatch.from_arrays(arrays, schema=schema)


def record_batch_to_segments(batch) -> Tuple[mc.HostSegments, Dict[str, np.ndarray]]:
    """The segment columns of a record batch / table with (a superset of) QUERY_COMPRESSED_SCHEMA as a HostSegments,
    plus the remaining columns (field_column, tags) as numpy arrays."""
    names = batch.schema.names
    missing = [f.name for f in query_compressed_schema() if f.name not in names and f.name != "error"]
    if missing:
        raise ValueError("not a batch of compressed segments, missing: " + ", ".join(missing))

    def numeric(name, arrow_type, dtype):
        column = batch.column(name)
        if isinstance(column, pa.ChunkedArray):
            column = column.combine_chunks()
        if column.null_count:
            raise ValueError(name + " is not nullable")
        return np.ascontiguousarray(column.cast(arrow_type).to_numpy(zero_copy_only=False), dtype=dtype)

    cols = dict(model_type_id=numeric("model_type_id", pa.int8(), np.int8),
                start_time=numeric("start_time", pa.int64(), np.int64), end_time=numeric("end_time", pa.int64(), np.int64),
                min_value=numeric("min_value", VALUE, np.float32), max_value=numeric("max_value", VALUE, np.float32))
    for name in _BINARY_COLUMNS:
        cols[name + "_off"], cols[name + "_data"] = _offsets_and_data(batch.column(name))
    extra = {}
    for name in names:
        if name in cols or name in _BINARY_COLUMNS or name == "error":
            continue
        column = batch.column(name)
        if pa.types.is_string_view(column.type) or pa.types.is_string(column.type) or pa.types.is_large_string(column.type):
            extra[name] = np.asarray(column.cast(pa.string()).to_pylist(), dtype=object)
        else:
            extra[name] = np.asarray(column.to_numpy(zero_copy_only=False) if not isinstance(column, pa.ChunkedArray) else column.combine_chunks().to_numpy(zero_copy_only=False))
    return mc.HostSegments(**cols), extra


# ---- Apache Parquet ----------------------------------------------------------------------------------------------------

def write_record_batch_to_apache_parquet_file(file_path: str, record_batch: pa.RecordBatch,
                                              sorting_columns: Optional[Sequence[Tuple[str, bool]]] = None) -> None:
    """lib.rs:216-261: `.parquet` extension required; 16 KiB data pages, row groups of 65 536 rows, PLAIN encoding, ZSTD
    at its default level, no dictionary, no statistics, no bloom filter.  sorting_columns: (name, descending) pairs."""
    if not str(file_path).endswith(".parquet"):
        raise ValueError("Apache Parquet file at path does not have the extension '.parquet'.")  # lib.rs:238-243
    sorting = None
    if sorting_columns:
        sorting = [pq.SortingColumn(record_batch.schema.get_field_index(name), descending=descending) for name, descending in sorting_columns]
    writer = pq.ParquetWriter(file_path, record_batch.schema, compression="zstd", use_dictionary=False, write_statistics=False,
                              data_page_size=16384, column_encoding={f.name: "PLAIN" for f in record_batch.schema},
                              sorting_columns=sorting)
    try:
        writer.write_table(pa.Table.from_batches([record_batch], record_batch.schema), row_group_size=65536)
    finally:
        writer.close()


def read_record_batches_from_apache_parquet_file(file_path: str) -> Iterator[pa.RecordBatch]:
    """lib.rs:173-210: one record batch per row group, binary columns as BinaryView like the reference's reader."""
    try:
        parquet_file = pq.ParquetFile(file_path)
    except (pa.ArrowInvalid, OSError) as error:
        raise ValueError("not an Apache Parquet file: %s" % error)
    for group in range(parquet_file.num_row_groups):
        table = parquet_file.read_row_group(group)
        for batch in table.combine_chunks().to_batches():
            yield batch
    parquet_file.close()


def write_segments(file_path: str, segments, field_column: Optional[int] = None, tags: Optional[Dict[str, object]] = None) -> None:
    """Segments of one compress call as a file the reference's reader accepts, sorted by start_time as the storage
    layer writes them."""
    write_record_batch_to_apache_parquet_file(file_path, segments_to_record_batch(segments, field_column, tags), [("start_time", False)])


def read_segments(file_path: str) -> Iterator[Tuple[mc.HostSegments, Dict[str, np.ndarray]]]:
    """(segments, extra columns) per row group: the input of operators.GridStream / the accumulators."""
    for batch in read_record_batches_from_apache_parquet_file(file_path):
        yield record_batch_to_segments(batch)


# ---- Arrow IPC ---------------------------------------------------------------------------------------------------------

def grid_record_batch(timestamps: np.ndarray, values: np.ndarray, tags: Sequence[np.ndarray] = (), tag_names: Sequence[str] = ()) -> pa.RecordBatch:
    """One output batch of GridStream as an Arrow record batch with GRID_SCHEMA + tags; the two numeric columns are
    zero-copy views of the numpy arrays."""
    if len(tags) != len(tag_names):
        raise ValueError("one name per tag column")
    arrays = [pa.array(np.ascontiguousarray(timestamps, np.int64), pa.int64()).cast(TIMESTAMP), pa.array(np.ascontiguousarray(values, np.float32), VALUE)]
    arrays += [pa.array(list(t), pa.string()).cast(pa.string_view()) for t in tags]
    return pa.RecordBatch.from_arrays(arrays, schema=grid_schema(tag_names))


def send_query_result(batches: Iterable[tuple], tag_names: Sequence[str] = (), sink=None) -> bytes:
    """remote.rs:169-211: the schema as one IPC message, then every record batch as one IPC message with the default
    IpcWriteOptions (no compression, metadata version 5).  `batches` are GridStream outputs (timestamps, values, tags...).
    Returns the stream bytes (or writes to `sink`)."""
    out = sink if sink is not None else io.BytesIO()
    schema = grid_schema(tag_names)
    with pa.ipc.new_stream(out, schema, options=pa.ipc.IpcWriteOptions()) as writer:
        for batch in batches:
            writer.write_batch(grid_record_batch(batch[0], batch[1], batch[2:], tag_names))
    return out.getvalue() if sink is None else b""


def read_query_result(stream: bytes) -> List[pa.RecordBatch]:
    """What a Flight client does with the messages of do_get."""
    with pa.ipc.open_stream(stream) as reader:
        return list(reader)
