"""GPU parity tests added in round 2: the device's own rewrite position for every f32, caller offsets that run
backwards or out of range, and two BASELINE.json configs at their exact / slab size against the oracle."""
import ctypes

import numpy as np
import pytest

from modelardb_rs_b200 import _native
from modelardb_rs_b200 import compression as mc
from modelardb_rs_b200 import operators as ops
from modelardb_rs_b200 import synthetic as syn
from tests.parity_cases import assert_f32_bits_equal, assert_segments_equal

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    return mc.Context(0)


def test_rewrite_position_on_the_device(oracle, ctx):
    """`23 - floor(|log2(x)|) as i32` (macaque_v.rs:185) decides which mantissa bits of every lossy MacaqueV value are
    cleared.  The reference calls libm's log2f, the device CUDA's f64 log2 rounded once (mdb_device.cuh:
    rewrite_position).  Both are step functions of the input: the device's steps over EVERY bit pattern from +0 to +inf
    must be the reference's, pattern for pattern (tests/test_oracle_log2.py is the host half of this check)."""
    want_bits, want_pos = oracle.rewrite_position_steps(0)
    cap = 4096
    bits = np.empty(cap, np.uint32)
    pos = np.empty(cap, np.int32)
    import ctypes as C
    n = C.c_uint32()
    _native.check(_native.lib().mdbcu_debug_rewrite_position_steps(ctx._h, 0, 0x7F800000, bits.ctypes.data, pos.ctypes.data, cap, C.byref(n)))
    assert n.value <= cap
    order = np.argsort(bits[: n.value])
    got_bits, got_pos = bits[: n.value][order], pos[: n.value][order]
    assert np.array_equal(got_bits, want_bits), (len(got_bits), len(want_bits))
    assert np.array_equal(got_pos, want_pos)


def _segments(oracle, n_series=4, n=3000, kind="sine", eb=(2, 1.0)):
    ts, vals, off = syn.multi_series(n_series, n, 11, kind)
    seg = oracle.compress(ts, vals, off, eb=eb)
    return seg, mc.HostSegments(unit_seg_off=seg.unit_seg_off, **{c: getattr(seg, c) for c in mc._COLUMNS})


@pytest.mark.parametrize("column", ["timestamps_off", "values_off", "residuals_off"])
@pytest.mark.parametrize("space", ["host", "device"])
def test_bad_offsets_fail_instead_of_reading_out_of_bounds(oracle, ctx, column, space):
    """The three offset columns are caller data.  Offsets that run backwards or point past the column's last offset make
    the call fail with the row's index; nothing is decoded from a length derived from them, and the context (and the
    process: no sticky CUDA error) stays usable."""
    import torch
    seg, host = _segments(oracle, kind="walk", eb=(2, 1.0))  # rows with values and residual bytes
    S = len(host)
    assert S > 8

    def run(bad_host):
        if space == "host":
            batch = bad_host
        else:
            cols = {}
            for c in mc._COLUMNS:
                a = getattr(bad_host, c)
                cols[c] = torch.from_numpy(a.view(np.int64) if a.dtype == np.uint64 else a).cuda()
            batch = mc.DeviceSegments(**cols)
        with pytest.raises(mc.ModelarDbCudaError, match="malformed segment row"):
            mc.grid(batch, ctx=ctx)
        with pytest.raises(mc.ModelarDbCudaError, match="malformed segment row"):
            mc.segment_sums(batch, ctx)
        with pytest.raises(mc.ModelarDbCudaError, match="malformed segment row"):
            mc.aggregate(batch, None, ctx)

    off = getattr(host, column).copy()
    # (1) backwards: two interior offsets swapped (only when they differ)
    k = next(i for i in range(1, S - 1) if off[i] != off[i + 1])
    swapped = off.copy()
    swapped[k], swapped[k + 1] = swapped[k + 1], swapped[k]
    # (2) an interior offset far past the end of the data
    huge = off.copy()
    huge[S // 2] = np.uint64(1) << np.uint64(40)
    for bad_off in (swapped, huge):
        cols = {c: getattr(host, c) for c in mc._COLUMNS}
        cols[column] = bad_off
        run(mc.HostSegments(unit_seg_off=host.unit_seg_off, **cols))
    # the context still works
    gts, gval = mc.grid(host, ctx=ctx)
    wts, wval, _ = oracle.grid(seg)
    assert np.array_equal(gts, wts)
    assert_f32_bits_equal(gval, wval, "grid after failed calls")


def test_config1_full_size_on_the_gpu(oracle, ctx):
    """BASELINE.json configs[0] at its exact size through the CUDA library: one series of 2^20 points (sine + noise,
    seed 1), lossless compress bit-identical to the oracle, then SELECT SUM / AVG through the model accumulators
    (crates/modelardb_embedded/src/operations/data_folder.rs:214-217, model_simple_aggregates.rs:473-618)."""
    n = 1 << 20
    ts = syn.regular_timestamps(n)
    vals = syn.sine_noise(n, 1)
    want = oracle.compress(ts, vals, eb=(0, 0.0))
    got = mc.try_compress_univariate_time_series(ts, vals, mc.Lossless, ctx)
    assert_segments_equal(got, want, "configs[0] compress")
    gts, gval = mc.grid(got, ctx=ctx)
    assert np.array_equal(gts, ts)
    assert_f32_bits_equal(gval, vals, "configs[0] grid")  # lossless: the input, bit for bit
    assert_f32_bits_equal(mc.segment_sums(got, ctx), oracle.segment_sums(want), "configs[0] row sums", nan_payload_matters=False)
    total, average = ops.ModelSumAccumulator(), ops.ModelAvgAccumulator()
    for lo in range(0, len(got), 1000):  # DataFusion feeds the accumulators batch by batch
        part = got.slice(lo, min(len(got), lo + 1000))
        total.update_batch(part)
        average.update_batch(part)
    (sum_state,), (count_state, avg_sum_state) = total.state(), average.state()
    assert count_state == n
    want_sum = 0.0
    for s in oracle.segment_sums(want):  # the reference's left fold of f32 row sums into f64
        want_sum += float(s)
    assert abs(sum_state - want_sum) <= 1e-12 * abs(want_sum) and abs(avg_sum_state - want_sum) <= 1e-12 * abs(want_sum)
    exact = float(vals.astype(np.float64).sum())
    assert abs(sum_state - exact) <= 1e-5 * abs(exact)  # integration_test.rs:1128-1171: 0.001 %


def test_config2_slab_of_1000_series_matches_oracle(oracle, ctx):
    """BASELINE.json configs[1] shape at a tenth of the series length: 1000 series x 10^5 points, 1 % relative bound,
    compress + full grid + GROUP BY series, every column and every point against the oracle."""
    n_series, n = 1000, 100_000
    ts, vals, off = syn.multi_series(n_series, n, 2, "sine")
    eb = (2, 1.0)
    want = oracle.compress(ts, vals, off, eb=eb, n_threads=16)
    seg = mc.compress(ts, vals, off, mc.ErrorBound(*eb), ctx)
    got = seg.to_host()
    assert_segments_equal(got, want, "configs[1] slab")
    wts, wval, _ = oracle.grid(want, n_threads=16)
    gts, gval = mc.grid(got, ctx=ctx)
    assert np.array_equal(gts, wts) and np.array_equal(gts, ts)
    assert_f32_bits_equal(gval, wval, "configs[1] slab grid")
    wc, wmn, wmx, wsm = oracle.aggregate(want, want.unit_seg_off, n_threads=16)
    gc, gmn, gmx, gsm = mc.aggregate(got, want.unit_seg_off, ctx)
    assert np.array_equal(gc, wc) and (gc == n).all()
    assert_f32_bits_equal(gmn, wmn)
    assert_f32_bits_equal(gmx, wmx)
    assert (np.abs(gsm - wsm) <= 1e-12 * np.abs(wsm)).all()
    rel = np.abs((vals.astype(np.float64) - gval.astype(np.float64)) / vals.astype(np.float64)) * 100.0
    assert float(rel.max()) <= 1.0  # within the user's bound of the raw input (compression.rs:914-928)
    seg.free()


@pytest.mark.parametrize("tile_kernel", ["search", "scan", "tma"])
@pytest.mark.parametrize("kind,eb", [("sine", (2, 1.0)), ("mixed", (2, 5.0)), ("walk", (2, 1.0)), ("sine", (0, 0.0)), ("sine", (2, 0.02))])
def test_tile_kernels_match_oracle(oracle, kind, eb, tile_kernel):
    """The three tile kernels of K2 -- k_grid_tile_search (rows of a tile by binary search per quad; the default), k_grid_tile
    (head flags + block-wide max-scan) and k_grid_tile_tma (tiles staged in shared memory and stored by the TMA engine,
    metadata prefetched with cp.async) -- write the same points as the oracle: PMC-Mean / Swing rows of every length (a 0.02 %
    bound gives hundreds of short rows per tile), residual tails and MacaqueV rows that the serial kernels overwrite
    afterwards, a last tile that is not a multiple of four points, and -- device space -- outputs that are not 16-byte
    aligned."""
    import torch
    ctx = mc.Context(0)
    ctx.set_option("grid_tma_stores", 1 if tile_kernel == "tma" else 0)
    ctx.set_option("grid_tile_scan", 1 if tile_kernel == "scan" else 0)
    ts, vals, off = syn.multi_series(7, 30_001, 19, kind)
    want = oracle.compress(ts, vals, off, eb=eb, n_threads=8)
    wts, wval, _ = oracle.grid(want, n_threads=8)
    host = mc.HostSegments(unit_seg_off=want.unit_seg_off, **{c: getattr(want, c) for c in mc._COLUMNS})
    gts, gval = mc.grid(host, ctx=ctx)
    assert np.array_equal(gts, wts)
    assert_f32_bits_equal(gval, wval, f"tma tile {kind} {eb}")
    seg = mc.compress(ts, vals, off, mc.ErrorBound(*eb), ctx)
    n = len(wts)
    for shift in (0, 1):  # shift 1: the outputs start 8 / 4 bytes past a 16-byte boundary
        tbuf = torch.zeros(n + 2, dtype=torch.int64, device="cuda:0")
        vbuf = torch.zeros(n + 4, dtype=torch.float32, device="cuda:0")
        dts, dval = mc.grid(seg, tbuf[shift:shift + n], vbuf[shift:shift + n], ctx)
        assert np.array_equal(dts.cpu().numpy(), wts), shift
        assert_f32_bits_equal(dval.cpu().numpy(), wval, f"tma tile device {kind} {eb} shift {shift}")
    seg.free()
    ctx.close()


@pytest.mark.parametrize("eb", [(0, 0.0), (2, 1e-4), (1, 0.05)], ids=["lossless", "rel1e-4", "abs0.05"])
def test_block_per_row_macaque_decoder_matches_oracle(oracle, eb):
    """k_macaque_block: a whole block decodes a very long MacaqueV row, 16 stretches of 256 fixed-width slots at a time, each
    stretch proven by its own flag bits; where a run breaks (a repeated value, a window that widens) one warp walks on the
    ordinary way.  The row length from which it is used is lowered ("block_row_min") so that rows of every length and kind
    of break go through it: grid values and per-row f32 sums (one addition chain in stream order) are the oracle's bits."""
    ctx = mc.Context(0)
    ctx.set_option("block_row_min", 512)
    rng = np.random.default_rng(5)
    units = []
    for u, n in enumerate([513, 4096, 4097, 9000, 40_000, 70_001]):
        x = (100.0 + np.cumsum(rng.standard_normal(n))).astype(np.float32)
        if u % 2:  # breaks: repeated values (`10` codes), an outlier that widens the window, a sign change
            x[n // 3: n // 3 + 5] = x[n // 3]
            x[n // 2] = np.float32(-3.0e30)
            x[2 * n // 3] = np.float32(1e-30)
        units.append(x)
    vals = np.concatenate(units)
    ts = np.concatenate([syn.regular_timestamps(len(u)) for u in units])
    off = np.concatenate([[0], np.cumsum([len(u) for u in units])]).astype(np.uint64)
    want = oracle.compress(ts, vals, off, eb=eb, n_threads=4)
    assert int(np.count_nonzero((want.model_type_id == 2) & (np.diff(want.values_off) > 1500))) >= 4  # long MacaqueV rows exist
    host = mc.HostSegments(unit_seg_off=want.unit_seg_off, **{c: getattr(want, c) for c in mc._COLUMNS})
    wts, wval, _ = oracle.grid(want, n_threads=4)
    gts, gval = mc.grid(host, ctx=ctx)
    assert np.array_equal(gts, wts)
    assert_f32_bits_equal(gval, wval, f"block decoder grid {eb}")
    assert_f32_bits_equal(mc.segment_sums(host, ctx), oracle.segment_sums(want, n_threads=4), f"block decoder sums {eb}", nan_payload_matters=False)
    gc, gmn, gmx, gsm = mc.aggregate(host, want.unit_seg_off, ctx)
    wc, wmn, wmx, wsm = oracle.aggregate(want, want.unit_seg_off, n_threads=4)
    assert np.array_equal(gc, wc)
    ok = np.isnan(wsm) & np.isnan(gsm) | (np.abs(gsm - wsm) <= 1e-12 * np.abs(wsm)) | (gsm == wsm)
    assert ok.all()
    ctx.close()


# ---- the time predicate inside the grid call (mdbcu_grid_range) -----------------------------------------------------------

def _range_cases(ts):
    t0, t1 = int(ts.min()), int(ts.max())
    span = t1 - t0
    return [(t0, t1), (t0 - 10, t1 + 10), (t0 + span // 3, t0 + span // 3 + span // 7), (t0 + 5, t0 + 5), (t0 + span // 2, t0 + span // 2 + 1),
            (t1, t1 + 100), (t0 - 100, t0), (t1 + 1, t1 + 50), (t0 - 50, t0 - 1), (t0 + 10, t0 + 9), (-(2 ** 63), 2 ** 63 - 1)]


@pytest.mark.parametrize("kind,eb,irregular", [("sine", (2, 1.0), False), ("mixed", (2, 5.0), False), ("walk", (0, 0.0), False),
                                              ("mixed", (1, 0.5), True), ("walk", (2, 1.0), True)])
def test_grid_range_equals_grid_then_filter(oracle, kind, eb, irregular):
    """mdbcu_grid_range returns exactly the points the reference keeps when it reconstructs everything and then prunes by
    `t_lo <= timestamp <= t_hi` (grid_exec.rs:366-387), in the same order, with the per-row counts -- for PMC-Mean / Swing /
    MacaqueV rows, residual tails, regular and irregular timestamps, host and device space, ranges that cut rows, hit a
    single point, cover everything or nothing."""
    import torch
    ctx = mc.Context(0)
    ts, vals, off = syn.multi_series(5, 20_011, 23, kind, irregular=irregular)
    want = oracle.compress(ts, vals, off, eb=eb, n_threads=8)
    wts, wval, wpo = oracle.grid(want, n_threads=8)
    row_of_point = np.repeat(np.arange(len(want)), np.diff(wpo).astype(np.int64))
    host = mc.HostSegments(unit_seg_off=want.unit_seg_off, **{c: getattr(want, c) for c in mc._COLUMNS})
    seg = mc.compress(ts, vals, off, mc.ErrorBound(*eb), ctx)
    for lo, hi in _range_cases(ts):
        keep = (wts >= lo) & (wts <= hi)
        counts = np.bincount(row_of_point[keep], minlength=len(want))
        gts, gval, gpo = mc.grid_range(host, lo, hi, ctx, with_point_off=True)
        assert np.array_equal(gts, wts[keep]), (lo, hi)
        assert_f32_bits_equal(gval, wval[keep], f"grid_range host {kind} {eb} [{lo}, {hi}]")
        assert np.array_equal(np.diff(gpo).astype(np.int64), counts), (lo, hi)
        dts, dval, dpo = mc.grid_range(seg, lo, hi, ctx, with_point_off=True)
        assert np.array_equal(dts.cpu().numpy(), wts[keep]), (lo, hi)
        assert_f32_bits_equal(dval.cpu().numpy(), wval[keep], f"grid_range device {kind} {eb} [{lo}, {hi}]")
        assert np.array_equal(np.diff(dpo.cpu().numpy()), counts), (lo, hi)
    n = ctypes.c_uint64()
    v = host.view()
    _native.check(_native.lib().mdbcu_grid_range(ctx._h, mc.HOST, ctypes.byref(v), int(wts[3]), int(wts[-3]), None, None, None, 0, ctypes.byref(n)))  # a count
    assert n.value == int(((wts >= wts[3]) & (wts <= wts[-3])).sum())
    with pytest.raises(mc.ModelarDbCudaError, match="capacity"):
        out_t, out_v = np.empty(2, np.int64), np.empty(2, np.float32)
        _native.check(_native.lib().mdbcu_grid_range(ctx._h, mc.HOST, ctypes.byref(v), int(wts.min()), int(wts.max()), None, out_t.ctypes.data,
                                                     out_v.ctypes.data, 2, ctypes.byref(n)))
    seg.free()
    ctx.close()


def test_grid_range_on_long_rows(oracle):
    """Rows of 10^5 MacaqueV values (copied by a block each) and whole-series PMC-Mean rows, cut in the middle."""
    ctx = mc.Context(0)
    n = 150_000
    ts = np.tile(syn.regular_timestamps(n), 3)
    rng = np.random.default_rng(5)
    vals = np.concatenate([(100 + np.cumsum(rng.standard_normal(n))).astype(np.float32), np.full(n, 7.5, np.float32),
                           (50 + 0.001 * np.arange(n)).astype(np.float32)])
    off = np.array([0, n, 2 * n, 3 * n], np.uint64)
    for eb in ((0, 0.0), (2, 1.0)):
        want = oracle.compress(ts, vals, off, eb=eb, n_threads=4)
        wts, wval, _ = oracle.grid(want, n_threads=4)
        host = mc.HostSegments(unit_seg_off=want.unit_seg_off, **{c: getattr(want, c) for c in mc._COLUMNS})
        lo, hi = int(ts[n // 3]), int(ts[n - n // 5])
        keep = (wts >= lo) & (wts <= hi)
        gts, gval = mc.grid_range(host, lo, hi, ctx)
        assert np.array_equal(gts, wts[keep])
        assert_f32_bits_equal(gval, wval[keep], f"grid_range long rows {eb}")
    ctx.close()


def test_grid_stream_with_the_time_predicate_on_the_device(oracle):
    """GridStream(time_range=..., device_time_clip=True): the same batches as evaluating the predicate on the host after
    reconstructing every point, which is what the reference does."""
    ctx = mc.Context(0)
    ts, vals, off = syn.multi_series(6, 9_000, 31, "mixed")
    seg = mc.compress(ts, vals, off, mc.ErrorBound(2, 5.0), ctx).to_host()
    tags = [np.array([f"s{u}" for u in np.repeat(np.arange(6), np.diff(seg.unit_seg_off).astype(np.int64))], object)]
    lo, hi = int(ts[1234]), int(ts[7000])

    def predicate(t, v):
        return (t >= lo) & (t <= hi)

    def batches(**kw):
        return list(ops.GridStream([(seg, tags)], batch_size=4096, n_tag_columns=1, predicate=predicate, ctx=ctx, **kw))

    a = batches()
    b = batches(time_range=(lo, hi), device_time_clip=True)
    assert len(a) == len(b) and len(a) > 1
    for x, y in zip(a, b):
        assert np.array_equal(x[0], y[0]) and np.array_equal(x[1].view(np.uint32), y[1].view(np.uint32)) and np.array_equal(x[2], y[2])
    ctx.close()


@pytest.mark.parametrize("overlap", [1, 0])
def test_a_unit_that_only_looks_regular_is_compressed_exactly(oracle, overlap):
    """The regularity check of the timestamps runs beside the screened chain kernel (mdbcu_compress, `overlap_regular_check`):
    the chains assume what the first interval and the last timestamp say.  A unit whose first interval and last timestamp are
    those of a regular unit but which has a shifted timestamp inside must still come out as the oracle's -- the chains are run
    again with the exact engine -- and so must the regular units beside it."""
    ctx = mc.Context(0)
    ctx.set_option("overlap_regular_check", overlap)
    n = 30_000
    ts, vals, off = syn.multi_series(4, n, 41, "sine")
    ts = ts.copy()
    ts[n + 12_345] += 1          # unit 1: one timestamp a millisecond late (first interval and last timestamp untouched)
    ts[3 * n + 2] += 3           # unit 3: the third point already
    for eb in ((2, 1.0), (1, 0.5)):
        want = oracle.compress(ts, vals, off, eb=eb, n_threads=4)
        got = mc.compress(ts, vals, off, mc.ErrorBound(*eb), ctx).to_host()
        assert_segments_equal(got, want, f"looks regular, overlap={overlap}, eb={eb}")
    ctx.close()
