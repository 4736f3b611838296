"""BASELINE.json configs[0] at FULL size on the CPU: one series of 2^20 points (sine + noise, seed 1),
lossless compress, then SUM and AVG through the model accumulators.  The kernel bodies run emulated (warp engine and
asynchronous scheduler for compress, the grid / sum bodies for the queries); the accumulators are the product's Python
operators with the oracle behind their one C-ABI call.  The GPU suite covers the same path through the CUDA library at
the sizes of tests/test_gpu_parity.py."""
import numpy as np

from modelardb_rs_b200 import compression as mc
from modelardb_rs_b200 import operators as ops
from modelardb_rs_b200 import synthetic as syn
from tests import emu_lib as emu
from tests.parity_cases import assert_f32_bits_equal, assert_segments_equal


def test_config1_compress_sum_avg(oracle, monkeypatch):
    n = 1 << 20
    ts = syn.regular_timestamps(n)
    vals = syn.sine_noise(n, 1)
    want = oracle.compress(ts, vals, eb=(0, 0.0))
    got = emu.compress(ts, vals, eb=(0, 0.0), chunk_len=65536, sched_seed=1, in_flight=8, engine=2)
    assert_segments_equal(got, want, "configs[0] compress")

    # lossless: the reconstructed points are the input, bit for bit
    gts, gval, _ = emu.grid(got)
    assert np.array_equal(gts, ts)
    assert_f32_bits_equal(gval, vals, "configs[0] grid")
    sums, counts = emu.segment_sums(got)
    assert_f32_bits_equal(sums, oracle.segment_sums(want), "configs[0] per-segment sums", nan_payload_matters=False)
    assert int(counts.sum()) == n

    # SELECT SUM / AVG via the model accumulators (model_simple_aggregates.rs:473-618)
    monkeypatch.setattr(mc, "aggregate", lambda host, group_off=None, ctx=None: oracle.aggregate(
        oracle.Segments(**{c: getattr(host, c) for c in mc._COLUMNS}), group_off))
    host = mc.HostSegments(**{c: getattr(got, c) for c in mc._COLUMNS})
    total, average = ops.ModelSumAccumulator(), ops.ModelAvgAccumulator()
    for lo in range(0, len(host), 1000):  # DataFusion feeds the accumulators batch by batch
        part = host.slice(lo, min(len(host), lo + 1000))
        total.update_batch(part)
        average.update_batch(part)
    (sum_state,), (count_state, avg_sum_state) = total.state(), average.state()
    exact = float(vals.astype(np.float64).sum())
    assert count_state == n
    # per-segment sums are f32 (models/mod.rs:129-184) added in f64: the reference's own test allows 0.001 % (integration_test.rs:1128-1171)
    assert abs(sum_state - exact) <= 1e-5 * abs(exact) and abs(avg_sum_state - exact) <= 1e-5 * abs(exact)
    assert abs(avg_sum_state / count_state - exact / n) <= 1e-5 * abs(exact / n)
    want_sum = float(np.sum(oracle.segment_sums(want).astype(np.float64)))
    assert abs(sum_state - want_sum) <= 1e-12 * abs(want_sum)
