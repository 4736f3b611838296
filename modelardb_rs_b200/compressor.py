"""Batching shim for the server's compressor thread.

The reference's compressor thread takes ONE finished buffer (<= 65 536 data points of one series, all its fields) at a
time from a channel, calls try_compress_univariate_time_series once per field and sends one CompressedSegmentBatch on
(crates/modelardb_server/src/storage/uncompressed_data_manager.rs:503-581).  One buffer per call would starve a GPU, so
this mirror of the loop drains every buffer that is already waiting in the channel -- up to a budget of data points --
and compresses all their (buffer, field) pairs with ONE mdbcu_compress; what leaves is exactly what the reference
sends: one CompressedSegmentBatch per buffer, in arrival order, Flush and Stop forwarded in their place in that order.
Nothing is delayed to fill a batch: a buffer that arrives alone is compressed alone.
"""
from __future__ import annotations

import queue
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import compression as mc

FLUSH = "Flush"  # Message::Flush
STOP = "Stop"    # Message::Stop


@dataclass
class UncompressedDataBuffer:
    """What compress_finished_buffer reads from a finished in-memory or on-disk buffer (:530-556): the data points of
    ONE series (one combination of tag values) for all fields of its table."""
    timestamps: np.ndarray                # int64[n], sorted
    field_columns: Sequence[np.ndarray]   # float32[n] per field
    field_column_indices: Sequence[int]   # index of each field in the table's schema
    error_bounds: Sequence[mc.ErrorBound] # one per field (time_series_table_metadata.error_bounds[index])
    tag_values: Sequence[str] = ()
    batch_ids: frozenset = frozenset()

    def __len__(self):
        return len(self.timestamps)


@dataclass
class CompressedSegmentBatch:
    """CompressedSegmentBatch::new(metadata, compressed_segments, batch_ids) (:571-577): one segment batch per field."""
    tag_values: Sequence[str]
    compressed_segments: List[Tuple[int, mc.HostSegments]]  # (field_column_index, segments), in field order
    batch_ids: frozenset = frozenset()


@dataclass
class CompressorMetrics:
    calls: int = 0      # batched compress calls issued
    buffers: int = 0    # buffers compressed
    units: int = 0      # (buffer, field) pairs compressed
    points: int = 0
    largest_call_buffers: int = 0


def compress_finished_buffers(buffers: Sequence[UncompressedDataBuffer], ctx: Optional[mc.Context] = None) -> List[CompressedSegmentBatch]:
    """compress_finished_buffer (:530-581) for several buffers at once: every (buffer, field) pair is one unit of one
    batched compress; the result is split back per buffer and field.  The C-ABI pairs one timestamp with every value,
    so a buffer's timestamps are repeated once per field."""
    if not buffers:
        return []
    ts_parts, val_parts, bounds, lens = [], [], [], []
    for b in buffers:
        if not (len(b.field_columns) == len(b.field_column_indices) == len(b.error_bounds)):
            raise ValueError("one index and one error bound per field column")
        for values in b.field_columns:
            if len(values) != len(b.timestamps):
                # the reference's expect(): "uncompressed_timestamps and uncompressed_values should have the same length."
                raise mc.ModelarDbCudaError("Uncompressed timestamps and uncompressed values have different lengths.")
            ts_parts.append(np.asarray(b.timestamps, np.int64))
            val_parts.append(np.asarray(values, np.float32))
            lens.append(len(values))
        bounds.extend(b.error_bounds)
    if not lens:
        return [CompressedSegmentBatch(b.tag_values, [], b.batch_ids) for b in buffers]
    unit_off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    seg = mc.compress(np.concatenate(ts_parts), np.concatenate(val_parts), unit_off, list(bounds), ctx)
    try:
        host = seg.to_host()
    finally:
        seg.free()
    rows = host.unit_seg_off
    out, u = [], 0
    for b in buffers:
        fields = []
        for index in b.field_column_indices:
            fields.append((int(index), host.slice(int(rows[u]), int(rows[u + 1]))))
            u += 1
        out.append(CompressedSegmentBatch(b.tag_values, fields, b.batch_ids))
    return out


def process_compressor_messages(uncompressed_data_receiver: "queue.Queue", compressed_data_sender: "queue.Queue",
                                ctx: Optional[mc.Context] = None, max_points_per_call: int = 1 << 28,
                                metrics: Optional[CompressorMetrics] = None) -> CompressorMetrics:
    """process_compressor_messages (:503-524).  Messages are UncompressedDataBuffer (Message::Data), FLUSH or STOP.
    Blocks for the next message like the reference's recv(); everything else that is ALREADY in the channel and is data
    joins the same compress call (at most max_points_per_call values per call, and at least one buffer)."""
    metrics = metrics or CompressorMetrics()
    held = None  # a message taken from the channel that did not fit the current call
    while True:
        message = held if held is not None else uncompressed_data_receiver.get()
        held = None
        if message == FLUSH:
            compressed_data_sender.put(FLUSH)
            continue
        if message == STOP:
            compressed_data_sender.put(STOP)
            return metrics
        pending = [message]
        values = len(message) * len(message.field_columns)
        while True:
            try:
                nxt = uncompressed_data_receiver.get_nowait()
            except queue.Empty:
                break
            if isinstance(nxt, str) or values + len(nxt) * len(nxt.field_columns) > max_points_per_call:
                held = nxt  # Flush / Stop keep their place behind the data before them; an over-budget buffer waits a call
                break
            pending.append(nxt)
            values += len(nxt) * len(nxt.field_columns)
        for batch in compress_finished_buffers(pending, ctx):
            compressed_data_sender.put(batch)
        metrics.calls += 1
        metrics.buffers += len(pending)
        metrics.units += sum(len(b.field_columns) for b in pending)
        metrics.points += values
        metrics.largest_call_buffers = max(metrics.largest_call_buffers, len(pending))
