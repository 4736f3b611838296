"""GPU parity tests: the CUDA kernels, called through the C-ABI (include/modelardb_cuda.h), against the
CPU oracle on the same seeded inputs.

Bars:
  * compress: every column of every segment row bit-identical to the oracle (boundaries, model ids,
    start/end times, min/max bit patterns, timestamps / values / residuals bytes);
  * grid: timestamps and values bit-identical to the oracle's grid;
  * aggregate: COUNT / MIN / MAX identical; per-row sums bit-identical; the f64 group SUM within
    TOL_SUM_REL of the oracle's sequential fold (the GPU folds rows in a fixed tree, not left to right).
"""
import numpy as np
import pytest

from modelardb_rs_b200 import compression as mc
from modelardb_rs_b200 import synthetic as syn
from tests.parity_cases import assert_f32_bits_equal, assert_segments_equal, small_cases

pytestmark = pytest.mark.gpu

# f64 accumulation of S f32 row sums in a different order: |delta| <= S * 2^-53 * sum|row sum|.
TOL_SUM_REL = 1e-12

CASES = small_cases()


def _ebs(ebs):
    return [mc.ErrorBound(k, v) for k, v in ebs]


@pytest.fixture(scope="module")
def ctx():
    return mc.Context(0)


def _to_host_segments(seg):
    """oracle.Segments -> HostSegments (same arrays)."""
    cols = {c: getattr(seg, c) for c in mc._COLUMNS}
    return mc.HostSegments(unit_seg_off=seg.unit_seg_off, **cols)


def _check_sum(got, want, where):
    scale = np.maximum(np.abs(want), 1e-300)
    both_nan = np.isnan(got) & np.isnan(want)
    same_inf = np.isinf(want) & (got == want)
    ok = both_nan | same_inf | (np.abs(got - want) <= TOL_SUM_REL * scale) | (got == want)
    assert ok.all(), f"{where}: {got[~ok][:5]} vs {want[~ok][:5]}"


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_compress_grid_aggregate_match_oracle_host_space(oracle, ctx, case):
    name, ts, vals, off, ebs = case
    want = oracle.compress(ts, vals, off, eb=ebs)
    seg = mc.compress(ts, vals, off, _ebs(ebs), ctx)
    got = seg.to_host()
    assert_segments_equal(got, want, name)

    wts, wval, woff = oracle.grid(want)
    goff, total = mc.grid_count(got, ctx)
    assert total == len(wts) and np.array_equal(goff, woff)
    gts, gval = mc.grid(got, ctx=ctx)
    assert np.array_equal(gts, wts), name
    assert_f32_bits_equal(gval, wval, name + " grid")
    assert np.array_equal(gts, ts)  # compression.rs:912

    assert_f32_bits_equal(mc.segment_sums(got, ctx), oracle.segment_sums(want), name + " segment sums", nan_payload_matters=False)
    for group_off in (None, want.unit_seg_off):
        wc, wmn, wmx, wsm = oracle.aggregate(want, group_off)
        gc, gmn, gmx, gsm = mc.aggregate(got, group_off, ctx)
        assert np.array_equal(gc, wc), name
        assert_f32_bits_equal(gmn, wmn, name + " min")
        assert_f32_bits_equal(gmx, wmx, name + " max")
        _check_sum(gsm, wsm, name + " sum")
    seg.free()


@pytest.mark.parametrize("case", [c for c in CASES if c[0].startswith(("sine-epoch", "walk-epoch", "ragged", "walk-irregular"))],
                         ids=lambda c: c[0])
def test_device_space_matches_host_space(oracle, ctx, case):
    """Same calls with torch CUDA tensors: nothing crosses PCIe, results identical."""
    import torch
    name, ts, vals, off, ebs = case
    dev = "cuda:0"
    d_ts, d_vals = torch.from_numpy(ts).to(dev), torch.from_numpy(vals).to(dev)
    d_off = torch.from_numpy(off.astype(np.int64)).to(dev)
    seg = mc.compress(d_ts, d_vals, d_off, _ebs(ebs), ctx)
    want = oracle.compress(ts, vals, off, eb=ebs)
    assert_segments_equal(seg.to_host(), want, name)
    gts, gval = mc.grid(seg, ctx=ctx)
    wts, wval, _ = oracle.grid(want)
    assert np.array_equal(gts.cpu().numpy(), wts)
    assert_f32_bits_equal(gval.cpu().numpy(), wval, name)
    uso = (seg.unit_seg_off_device_ptr(), seg.n_units)
    gc, gmn, gmx, gsm = mc.aggregate(seg, uso, ctx)
    wc, wmn, wmx, wsm = oracle.aggregate(want, want.unit_seg_off)
    assert np.array_equal(gc.cpu().numpy(), wc)
    assert_f32_bits_equal(gmn.cpu().numpy(), wmn)
    assert_f32_bits_equal(gmx.cpu().numpy(), wmx)
    _check_sum(gsm.cpu().numpy(), wsm, name)
    seg.free()


@pytest.mark.parametrize("engine", [1, 2, 3, 4, 5], ids=["thread-per-chain", "warp-per-chain", "warp-per-chain-async", "lane-per-chain", "screened"])
@pytest.mark.parametrize("chunk_len", [8, 64, 1000, 4096])
def test_parallel_segmentation_is_independent_of_chunk_length(oracle, chunk_len, engine):
    """The chunked, speculative, fixpoint-stitched segmentation (csrc/mdb_compress.cuh) must yield exactly
    the sequential chain's rows for every chunk length, including chunks far shorter than a segment, and
    with either fit engine (one thread, or 32 lanes cooperating on each fit: csrc/mdb_fit_warp.cuh)."""
    ctx = mc.Context(0)
    ctx.set_chunk_len(chunk_len)
    ctx.set_fit_engine(engine)
    for name, ts, vals, off, ebs in CASES:
        if len(ts) > 20_000 and chunk_len < 64:
            continue
        want = oracle.compress(ts, vals, off, eb=ebs)
        seg = mc.compress(ts, vals, off, _ebs(ebs), ctx)
        assert_segments_equal(seg.to_host(), want, f"{name} chunk_len={chunk_len}")
        seg.free()
    # models much longer than a chunk: budgeted speculation, cut-short chains, skipped chunks
    rng = np.random.default_rng(3)
    vals = np.concatenate([np.full(5000, 3.25, np.float32), rng.uniform(0, 1, 37).astype(np.float32), np.full(3000, -7.5, np.float32),
                           (0.5 * np.arange(4000)).astype(np.float32), rng.uniform(0, 1, 300).astype(np.float32),
                           np.full(2500, 1.0, np.float32)])
    ts = syn.regular_timestamps(len(vals))
    for eb in ((0, 0.0), (2, 1.0), (1, 0.1)):
        want = oracle.compress(ts, vals, eb=eb)
        seg = mc.compress(ts, vals, None, mc.ErrorBound(*eb), ctx)
        assert_segments_equal(seg.to_host(), want, f"long models eb={eb} chunk_len={chunk_len}")
        if chunk_len <= 1000 and engine < 3:  # (the asynchronous scheduler has no rounds)
            assert ctx.last_compress_rounds >= 2
        seg.free()
    ctx.close()


def test_reference_known_answer_segment(ctx):
    # compression.rs:932-978 (through the public function: 5 points never reach a model -> one MacaqueV row)
    got = mc.try_compress_univariate_time_series(np.arange(100, 600, 100), np.array([73.0, 37.0, 37.0, 37.0, 73.0], np.float32),
                                                 mc.Lossless, ctx)
    assert len(got) == 1
    row = got.row(0)
    assert row["model_type_id"] == mc.MACAQUE_V_ID and (row["start_time"], row["end_time"]) == (100, 500)
    assert row["timestamps"] == bytes([5]) and row["residuals"] == b""
    assert float(row["min_value"]) == 37.0 and float(row["max_value"]) == 73.0
    assert list(row["values"]) == [66, 146, 0, 0, 208, 60, 58, 67]


def test_swing_reconstructs_linear_sequence_exactly(ctx):
    # swing.rs:717-798
    for reverse in (False, True):
        values = np.arange(42, 4201, 42).astype(np.float32)
        if reverse:
            values = values[::-1].copy()
        ts = 1658671178037 + 1000 * np.arange(len(values), dtype=np.int64)
        seg = mc.try_compress_univariate_time_series(ts, values, mc.Lossless, ctx)
        assert len(seg) == 1 and seg.model_type_id[0] == mc.SWING_ID
        gts, gval = mc.grid(seg, ctx=ctx)
        assert np.array_equal(gts, ts) and np.array_equal(gval, values)


def test_row_wise_api(oracle, ctx):
    # models/mod.rs:408-416 and :432-464
    assert mc.len_(1658671178037, 1658671178037, b"", ctx) == 1
    assert mc.len_(1658671178037, 1658671187047, bytes([10]), ctx) == 10
    ts, val = mc.grid_row(mc.PMC_MEAN_ID, 100, 500, bytes([5]), 10.0, 10.0, b"", b"", ctx)
    assert list(ts) == [100, 200, 300, 400, 500] and list(val) == [10.0] * 5
    # swing.rs:668-677: sum(START, END, [], first, last, 0) == first + last
    assert mc.sum_(mc.SWING_ID, 1658671178037, 1658671179037, b"", 3.0, 8.0, b"", b"", ctx) == np.float32(11.0)


def test_empty_inputs(ctx):
    # compression.rs:208-211 and a batch with no units / no rows
    seg = mc.try_compress_univariate_time_series(np.zeros(0, np.int64), np.zeros(0, np.float32), mc.Lossless, ctx)
    assert len(seg) == 0
    ts, val = mc.grid(seg, ctx=ctx)
    assert len(ts) == 0 and len(val) == 0
    c, mn, mx, sm = mc.aggregate(seg, None, ctx)
    assert c[0] == 0 and mn[0] == np.finfo(np.float32).max and mx[0] == -np.finfo(np.float32).max and sm[0] == 0.0
    seg2 = mc.compress(np.zeros(0, np.int64), np.zeros(0, np.float32), np.zeros(1, np.uint64), [], ctx)
    assert len(seg2) == 0


def test_error_behaviour(ctx):
    with pytest.raises(mc.ModelarDbCudaError, match="different lengths"):  # compression.rs:202-206
        mc.compress(np.zeros(3, np.int64), np.zeros(2, np.float32), None, mc.Lossless, ctx)
    with pytest.raises(mc.ModelarDbCudaError, match="malformed unit"):     # invalid bound reaches the library
        mc.compress(np.zeros(3, np.int64), np.zeros(3, np.float32), None, mc.ErrorBound(2, 150.0), ctx)
    # rows the reference would panic on (models/mod.rs:237, types.rs:405) fail instead
    bad = mc._one_row(7, 0, 10, b"", 0.0, 0.0, b"", b"")
    with pytest.raises(mc.ModelarDbCudaError, match="malformed segment row 0"):
        mc.grid(bad, ctx=ctx)
    bad = mc._one_row(mc.SWING_ID, 0, 10, b"", 0.0, 0.0, b"\x01\x02\x03", b"")
    with pytest.raises(mc.ModelarDbCudaError, match="malformed segment row 0"):
        mc.segment_sums(bad, ctx)
    good = mc._one_row(mc.PMC_MEAN_ID, 100, 500, bytes([5]), 10.0, 10.0, b"", b"")
    with pytest.raises(mc.ModelarDbCudaError, match="capacity"):
        mc.grid(good, np.empty(3, np.int64), np.empty(3, np.float32), ctx)


def test_server_path_buffers_match_oracle(oracle, ctx):
    """The server cuts every series into <= 65 536-point buffers before compressing
    (modelardb_server/src/storage/mod.rs:58): same kernels, different unit_off."""
    n_series, n = 3, 150_000
    ts, vals, _ = syn.multi_series(n_series, n, 77, "sine")
    off = mc.split_into_buffers([n] * n_series)
    eb = (2, 1.0)
    want = oracle.compress(ts, vals, off, eb=eb, n_threads=8)
    got = mc.compress(ts, vals, off, mc.ErrorBound(*eb), ctx).to_host()
    assert_segments_equal(got, want, "server buffers")


@pytest.mark.parametrize("kind,eb", [("sine", (2, 1.0)), ("walk", (0, 0.0)), ("sine", (2, 5.0)), ("walk", (2, 1.0))])
def test_medium_sizes_match_oracle(oracle, ctx, kind, eb):
    """BASELINE.json configs 2-5 at a size the oracle finishes in seconds (64 series x 50 000 points)."""
    ts, vals, off = syn.multi_series(64, 50_000, 123, kind)
    want = oracle.compress(ts, vals, off, eb=eb, n_threads=8)
    seg = mc.compress(ts, vals, off, mc.ErrorBound(*eb), ctx)
    got = seg.to_host()
    assert_segments_equal(got, want, f"{kind} {eb}")
    wts, wval, _ = oracle.grid(want, n_threads=8)
    gts, gval = mc.grid(got, ctx=ctx)
    assert np.array_equal(gts, wts)
    assert_f32_bits_equal(gval, wval, f"{kind} {eb} grid")
    wc, wmn, wmx, wsm = oracle.aggregate(want, want.unit_seg_off, n_threads=8)
    gc, gmn, gmx, gsm = mc.aggregate(got, want.unit_seg_off, ctx)
    assert np.array_equal(gc, wc)
    assert_f32_bits_equal(gmn, wmn)
    assert_f32_bits_equal(gmx, wmx)
    _check_sum(gsm, wsm, f"{kind} {eb}")
    # every reconstructed value is within the bound of the raw input (compression.rs:914-928)
    idx = np.random.default_rng(1).integers(0, len(vals), 20_000)
    for i in idx:
        assert oracle.is_value_within_error_bound(eb, vals[i], gval[i])
    seg.free()


def test_pageable_and_pinned_host_memory_give_the_same_bytes(oracle, ctx):
    """Host-space calls bounce pageable caller memory through a 2 x 32 MiB pinned ring and copy pinned memory
    directly: 9 M points (72 MB of timestamps in, 72 MB out) cross the ring halves several times each way."""
    import torch
    n_series, n = 9, 1_000_000
    ts, vals, off = syn.multi_series(n_series, n, 5, "sine")
    eb = (2, 1.0)
    want = oracle.compress(ts, vals, off, eb=eb, n_threads=8)
    wts, wval, _ = oracle.grid(want, n_threads=8)
    # pageable in, pageable out
    seg = mc.compress(ts, vals, off, mc.ErrorBound(*eb), ctx)
    got = seg.to_host()
    assert_segments_equal(got, want, "pageable")
    gts = np.full(len(ts), -1, np.int64)
    gval = np.full(len(ts), np.nan, np.float32)
    mc.grid(got, gts, gval, ctx)
    assert np.array_equal(gts, wts)
    assert_f32_bits_equal(gval, wval, "pageable grid")
    # pinned in, pinned out, and the zero-copy view of the library's host copy
    pts, pvals = torch.from_numpy(ts).pin_memory(), torch.from_numpy(vals).pin_memory()
    seg2 = mc.compress(pts.numpy(), pvals.numpy(), off, mc.ErrorBound(*eb), ctx)
    view = seg2.to_host(copy=False)
    assert_segments_equal(view, want, "pinned")
    pts_out = torch.empty(len(ts), dtype=torch.int64).pin_memory()
    pval_out = torch.empty(len(ts), dtype=torch.float32).pin_memory()
    mc.grid(view, pts_out.numpy(), pval_out.numpy(), ctx)
    assert np.array_equal(pts_out.numpy(), wts)
    assert_f32_bits_equal(pval_out.numpy(), wval, "pinned grid")
    gc, _, _, gsm = mc.aggregate(view, want.unit_seg_off, ctx)
    wc, _, _, wsm = oracle.aggregate(want, want.unit_seg_off, n_threads=8)
    assert np.array_equal(gc, wc)
    _check_sum(gsm, wsm, "pinned aggregate")
    del view
    seg.free()
    seg2.free()


@pytest.mark.parametrize("eb", [(0, 0.0), (1, 1e-3), (2, 1e-4), (2, 3.0)], ids=["lossless", "abs1e-3", "rel1e-4", "rel3"])
def test_long_macaque_v_rows_match_oracle(oracle, ctx, eb):
    """Rows of tens of thousands of MacaqueV values: the warp encoder (ballot loops for stored-value and window
    resets, scanned bit offsets, several drains of the shared stage) and the warp decoder (several refills of the
    stream stage), lossless and lossy, against the oracle's serial coder."""
    rng = np.random.default_rng(17)
    units = [rng.uniform(-1e3, 1e3, 50_000).astype(np.float32),                       # no model ever fits
             (100.0 + np.cumsum(rng.standard_normal(70_001))).astype(np.float32),     # random walk
             np.repeat(rng.uniform(-5, 5, 4_000).astype(np.float32), 9)[:33_333]]     # runs of 9 equal values: `10` codes, models between
    vals = np.concatenate(units)
    ts = np.concatenate([syn.regular_timestamps(len(u)) for u in units])
    off = np.concatenate([[0], np.cumsum([len(u) for u in units])]).astype(np.uint64)
    want = oracle.compress(ts, vals, off, eb=eb, n_threads=4)
    seg = mc.compress(ts, vals, off, mc.ErrorBound(*eb), ctx)
    got = seg.to_host()
    assert_segments_equal(got, want, f"long MacaqueV rows eb={eb}")
    wts, wval, _ = oracle.grid(want, n_threads=4)
    gts, gval = mc.grid(got, ctx=ctx)
    assert np.array_equal(gts, wts)
    assert_f32_bits_equal(gval, wval, f"long MacaqueV rows eb={eb} grid")
    wsum = oracle.segment_sums(want)
    gsum = mc.segment_sums(got, ctx)
    assert_f32_bits_equal(gsum, wsum, f"long MacaqueV rows eb={eb} sums", nan_payload_matters=False)
    seg.free()


@pytest.mark.parametrize("eb", [(0, 0.0), (1, 0.05), (2, 1e-4)], ids=["lossless", "abs0.05", "rel1e-4"])
def test_thousands_of_long_macaque_v_rows_match_oracle(oracle, monkeypatch, eb):
    """From tens of thousands of long MacaqueV rows in a batch on (24 576; lowered to 4096 here), grid and the aggregates
    give every row to one thread instead of one warp (k_grid_macaque_lanes / k_agg_macaque_lanes): 5000 series of a few
    hundred values, no model ever fits (some series hold runs of equal values so that models and residuals sit between
    the MacaqueV rows)."""
    ctx = mc.Context(0)
    ctx.set_option("lane_rows_min", 4096)
    rng = np.random.default_rng(23)
    lens = rng.integers(150, 400, 5000)
    units = []
    for u, n in enumerate(lens):
        x = rng.uniform(-1e3, 1e3, n).astype(np.float32)
        if u % 7 == 0:
            x[n // 3: n // 3 + 40] = x[n // 3]  # a constant stretch: a PMC-Mean model, residuals before the next row
        units.append(x)
    vals = np.concatenate(units)
    ts = np.concatenate([syn.regular_timestamps(len(u)) for u in units])
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    want = oracle.compress(ts, vals, off, eb=eb, n_threads=8)
    assert int(np.count_nonzero((want.model_type_id == 2) & (np.diff(want.values_off) > 200))) >= 4096
    seg = mc.compress(ts, vals, off, mc.ErrorBound(*eb), ctx)
    got = seg.to_host()
    assert_segments_equal(got, want, f"many MacaqueV rows eb={eb}")
    wts, wval, _ = oracle.grid(want, n_threads=8)
    gts, gval = mc.grid(got, ctx=ctx)
    assert np.array_equal(gts, wts)
    assert_f32_bits_equal(gval, wval, f"many MacaqueV rows eb={eb} grid")
    dts, dval = mc.grid(seg, ctx=ctx)  # device-resident segments and outputs
    assert np.array_equal(dts.cpu().numpy(), wts)
    assert_f32_bits_equal(dval.cpu().numpy(), wval, f"many MacaqueV rows eb={eb} device grid")
    assert_f32_bits_equal(mc.segment_sums(got, ctx), oracle.segment_sums(want, n_threads=8), f"many MacaqueV rows eb={eb} sums",
                          nan_payload_matters=False)
    gc, gmn, gmx, gsm = mc.aggregate(got, want.unit_seg_off, ctx)
    wc, wmn, wmx, wsm = oracle.aggregate(want, want.unit_seg_off, n_threads=8)
    assert np.array_equal(gc, wc) and gmn.tobytes() == wmn.tobytes() and gmx.tobytes() == wmx.tobytes()
    _check_sum(gsm, wsm, "many MacaqueV rows aggregate")
    seg.free()
    ctx.close()


def test_lane_and_warp_decoders_agree(oracle, monkeypatch):
    """The row count that switches between the two decoders is a tuning knob (mdbcu_context_set_option "lane_rows_min"): the
    same batch through both gives the same bits.  The tile kernel with per-thread stores and the one that hands whole tiles
    to the TMA engine are compared the same way ("grid_tma_stores")."""
    ts, vals, off = syn.multi_series(40, 3000, 77, "walk")
    want = oracle.compress(ts, vals, off, eb=(0, 0.0), n_threads=4)
    wts, wval, _ = oracle.grid(want, n_threads=4)
    host = mc.HostSegments(**{c: getattr(want, c) for c in mc._COLUMNS})
    for rows_min in (1, 1000000):
        c = mc.Context(0)
        c.set_option("lane_rows_min", rows_min)
        c.set_option("grid_tma_stores", 1 if rows_min == 1 else 0)
        with pytest.raises(mc.ModelarDbCudaError, match="unknown option"):
            c.set_option("no_such_knob", 1)
        gts, gval = mc.grid(host, ctx=c)
        assert np.array_equal(gts, wts)
        assert_f32_bits_equal(gval, wval, f"lane rows min {rows_min}")
        assert_f32_bits_equal(mc.segment_sums(host, c), oracle.segment_sums(want), f"sums, lane rows min {rows_min}", nan_payload_matters=False)
        c.close()


def test_contexts_on_threads_pipeline_independent_slabs(oracle):
    """One context per host thread (the e2e pattern of bench.py): results do not depend on what the others do."""
    import threading
    ts, vals, off = syn.multi_series(8, 200_000, 9, "walk")
    eb = (2, 1.0)
    want = oracle.compress(ts, vals, off, eb=eb, n_threads=8)
    wts, wval, _ = oracle.grid(want, n_threads=8)
    errors = []

    def work():
        try:
            c = mc.Context(0)
            for _ in range(3):
                seg = mc.compress(ts, vals, off, mc.ErrorBound(*eb), c)
                host = seg.to_host(copy=False)
                assert_segments_equal(host, want, "threaded")
                gts, gval = mc.grid(host, ctx=c)
                assert np.array_equal(gts, wts)
                assert_f32_bits_equal(gval, wval, "threaded grid")
                del host
                seg.free()
            c.close()
        except Exception as e:  # noqa: BLE001 -- reported by the main thread
            errors.append(e)

    threads = [threading.Thread(target=work) for _ in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


def test_full_size_properties_device(ctx):
    """Size-independent properties at a bench-sized slab (no oracle run): lossless round trip is exact,
    timestamps round-trip exactly, COUNT equals the number of points, grouped aggregates add up to the
    ungrouped ones, and compress is deterministic."""
    import torch
    dev = "cuda:0"
    n_series, n = 256, 200_000
    g = torch.Generator(device=dev).manual_seed(5)
    i = torch.arange(n, device=dev, dtype=torch.float64)
    phase = torch.rand(n_series, 1, device=dev, generator=g, dtype=torch.float64) * 6.28
    vals = (100.0 + 10.0 * torch.sin(2 * torch.pi * i / 1000.0 + phase)
            + 0.1 * torch.randn(n_series, n, device=dev, generator=g, dtype=torch.float64)).to(torch.float32).reshape(-1)
    ts = (syn.EPOCH_US + syn.STEP_US * torch.arange(n, device=dev, dtype=torch.int64)).repeat(n_series)
    off = torch.arange(n_series + 1, device=dev, dtype=torch.int64) * n
    for eb in (mc.Lossless, mc.ErrorBound.try_new_relative(1.0)):
        seg = mc.compress(ts, vals, off, eb, ctx)
        gts, gval = mc.grid(seg, ctx=ctx)
        assert torch.equal(gts, ts)
        if eb.kind == 0:
            assert torch.equal(gval.view(torch.int32), vals.view(torch.int32))
        else:
            rel = ((vals - gval) / vals).abs() * 100.0
            assert bool((rel <= 1.0).all())
        uso = (seg.unit_seg_off_device_ptr(), n_series)
        c, mn, mx, sm = mc.aggregate(seg, uso, ctx)
        assert bool((c == n).all())
        c1, mn1, mx1, sm1 = mc.aggregate(seg, None, ctx)
        assert int(c1[0]) == n_series * n
        assert float(mn1[0]) == float(mn.min()) and float(mx1[0]) == float(mx.max())
        assert abs(float(sm1[0]) - float(sm.sum())) <= 1e-9 * abs(float(sm1[0]))
        ref = gval.to(torch.float64).reshape(n_series, n).sum(1)
        # Model rows: within 0.001 % of the aggregate over data points (integration_test.rs:1184-1246).
        # Lossless noise is one MacaqueV row per series whose sum the reference accumulates sequentially in
        # f32 (macaque_v.rs:228-264): the bound there is the f32 reduction-order one, n * 2^-24 relative.
        tol = 1e-5 if eb.kind != 0 else n * 2.0 ** -24
        assert bool(((sm - ref).abs() <= tol * ref.abs()).all())
        again = mc.compress(ts, vals, off, eb, ctx)
        a, b = seg.to_host(), again.to_host()
        assert_segments_equal(a, b, "determinism")
        seg.free()
        again.free()
