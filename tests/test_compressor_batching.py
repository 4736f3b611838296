"""modelardb_rs_b200/compressor.py: the batching shim of the server's compressor thread
(uncompressed_data_manager.rs:503-581).  What leaves the shim must be what the reference sends: one
CompressedSegmentBatch per buffer with the segments try_compress_univariate_time_series gives for each field, in arrival
order, Flush / Stop in their place.  The CPU tests put the oracle behind the one C-ABI call; the gpu test runs the
CUDA library."""
import queue
import threading

import numpy as np
import pytest

from modelardb_rs_b200 import compression as mc
from modelardb_rs_b200 import compressor as comp
from modelardb_rs_b200 import synthetic as syn
from tests.parity_cases import assert_segments_equal


def _buffers(n_buffers=7, seed=3):
    rng = np.random.default_rng(seed)
    out = []
    for b in range(n_buffers):
        n = int(rng.integers(1, 3000)) if b != 2 else 65536
        ts = syn.regular_timestamps(n) + b * 10**9
        fields = [syn.sine_noise(n, seed + b), syn.random_walk(n, seed + 50 + b), np.full(n, 3.25, np.float32)]
        bounds = [mc.ErrorBound.try_new_relative(1.0), mc.ErrorBound.lossless(), mc.ErrorBound.try_new_absolute(0.5)]
        out.append(comp.UncompressedDataBuffer(ts, fields, [1, 2, 4], bounds, ["tag-%d" % b], frozenset([b])))
    return out


def _check_batches(oracle, buffers, batches):
    assert len(batches) == len(buffers)
    for buffer, batch in zip(buffers, batches):
        assert list(batch.tag_values) == list(buffer.tag_values) and batch.batch_ids == buffer.batch_ids
        assert [i for i, _ in batch.compressed_segments] == list(buffer.field_column_indices)
        for (_, got), values, bound in zip(batch.compressed_segments, buffer.field_columns, buffer.error_bounds):
            want = oracle.compress(buffer.timestamps, values, None, eb=(bound.kind, bound.value))
            assert_segments_equal(got, want, "buffer %s" % list(buffer.tag_values))


@pytest.fixture
def oracle_behind_compress(oracle, monkeypatch):
    calls = []

    class Owned:
        def __init__(self, seg):
            self.seg = seg

        def to_host(self, copy=True):
            return mc.HostSegments(unit_seg_off=self.seg.unit_seg_off, **{c: getattr(self.seg, c) for c in mc._COLUMNS})

        def free(self):
            pass

    def compress(timestamps, values, unit_off=None, error_bound=mc.Lossless, ctx=None):
        n_units = 1 if unit_off is None else len(unit_off) - 1
        kinds, vals = mc._bounds(error_bound, n_units)
        calls.append(n_units)
        return Owned(oracle.compress(timestamps, values, unit_off, eb=[(int(k), float(v)) for k, v in zip(kinds, vals)]))

    monkeypatch.setattr(mc, "compress", compress)
    return oracle, calls


def test_many_buffers_one_call(oracle_behind_compress):
    oracle, calls = oracle_behind_compress
    buffers = _buffers()
    batches = comp.compress_finished_buffers(buffers)
    assert calls == [3 * len(buffers)]  # every (buffer, field) pair is a unit of ONE call
    _check_batches(oracle, buffers, batches)
    assert comp.compress_finished_buffers([]) == []


def test_length_mismatch_is_the_references_error(oracle_behind_compress):
    b = _buffers(1)[0]
    b.field_columns[1] = b.field_columns[1][:-1]
    with pytest.raises(mc.ModelarDbCudaError, match="different lengths"):
        comp.compress_finished_buffers([b])


def test_message_loop_keeps_order_and_forwards_flush_and_stop(oracle_behind_compress):
    oracle, calls = oracle_behind_compress
    buffers = _buffers(9)
    receiver, sender = queue.Queue(), queue.Queue()
    for m in buffers[:4] + [comp.FLUSH] + buffers[4:6] + [comp.FLUSH, comp.FLUSH] + buffers[6:] + [comp.STOP]:
        receiver.put(m)
    metrics = comp.process_compressor_messages(receiver, sender)
    out = []
    while not sender.empty():
        out.append(sender.get())
    kinds = ["data" if isinstance(m, comp.CompressedSegmentBatch) else m for m in out]
    assert kinds == ["data"] * 4 + [comp.FLUSH] + ["data"] * 2 + [comp.FLUSH, comp.FLUSH] + ["data"] * 3 + [comp.STOP]
    _check_batches(oracle, buffers, [m for m in out if isinstance(m, comp.CompressedSegmentBatch)])
    # everything that was waiting between two control messages went into one call
    assert calls == [12, 6, 9] and metrics.calls == 3 and metrics.buffers == 9 and metrics.largest_call_buffers == 4
    assert metrics.points == sum(3 * len(b) for b in buffers)
    assert receiver.empty()


def test_point_budget_splits_calls_but_not_results(oracle_behind_compress):
    oracle, calls = oracle_behind_compress
    buffers = _buffers(6)
    receiver, sender = queue.Queue(), queue.Queue()
    for m in buffers + [comp.STOP]:
        receiver.put(m)
    metrics = comp.process_compressor_messages(receiver, sender, max_points_per_call=3 * 4000)
    out = [sender.get() for _ in range(len(buffers))]
    assert sender.get() == comp.STOP
    _check_batches(oracle, buffers, out)
    assert metrics.calls > 1 and sum(calls) == 3 * len(buffers)
    assert 3 in calls  # the 65 536-point buffer exceeds the budget by itself and is still compressed, alone


def test_a_lone_buffer_is_not_held_back(oracle_behind_compress):
    """Nothing waits for a batch to fill: with a slow producer every buffer leaves as soon as it has been compressed."""
    oracle, calls = oracle_behind_compress
    buffers = _buffers(3)
    receiver, sender = queue.Queue(), queue.Queue()
    worker = threading.Thread(target=comp.process_compressor_messages, args=(receiver, sender))
    worker.start()
    got = []
    for b in buffers:
        receiver.put(b)
        got.append(sender.get(timeout=60))  # arrives without any further message being sent
    receiver.put(comp.STOP)
    assert sender.get(timeout=60) == comp.STOP
    worker.join(timeout=60)
    assert not worker.is_alive() and calls == [3, 3, 3]
    _check_batches(oracle, buffers, got)


@pytest.mark.gpu
def test_compressor_loop_on_the_device(oracle):
    buffers = _buffers(12, seed=11)
    receiver, sender = queue.Queue(), queue.Queue()
    for m in buffers[:7] + [comp.FLUSH] + buffers[7:] + [comp.STOP]:
        receiver.put(m)
    metrics = comp.process_compressor_messages(receiver, sender, mc.Context(0))
    out = []
    while not sender.empty():
        out.append(sender.get())
    assert out[7] == comp.FLUSH and out[-1] == comp.STOP and metrics.calls == 2
    _check_batches(oracle, buffers, [m for m in out if isinstance(m, comp.CompressedSegmentBatch)])
