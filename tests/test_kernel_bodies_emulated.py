"""CPU check of the per-thread kernel bodies (csrc/mdb_*.cuh compiled for the host, tests/emu/emu.cc)
against the oracle.  This is the fast inner loop on the GPU-less build box; the bit-exactness CLAIM is
made by the `-m gpu` tests, which run the real kernels through the C-ABI on the same cases."""
import numpy as np
import pytest

from modelardb_rs_b200 import synthetic as syn
from tests import emu_lib as emu
from tests.parity_cases import assert_f32_bits_equal, assert_segments_equal, small_cases

CASES = small_cases()


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_compress_grid_aggregate_bodies_match_oracle(oracle, case):
    name, ts, vals, off, ebs = case
    want = oracle.compress(ts, vals, off, eb=ebs)
    got = emu.compress(ts, vals, off, eb=ebs)
    assert_segments_equal(got, want, name)

    wts, wval, woff = oracle.grid(want)
    gts, gval, goff = emu.grid(want)
    assert np.array_equal(goff, woff)
    assert np.array_equal(gts, wts)
    assert_f32_bits_equal(gval, wval, name + " grid")
    assert np.array_equal(gts, ts)  # timestamps always round-trip exactly (compression.rs:912)

    wsum = oracle.segment_sums(want)
    gsum, gcount = emu.segment_sums(want)
    assert_f32_bits_equal(gsum, wsum, name + " segment sums", nan_payload_matters=False)
    assert np.array_equal(gcount, np.diff(woff))

    wc, wmn, wmx, wsm = oracle.aggregate(want, want.unit_seg_off)
    gc, gmn, gmx, gsm = emu.aggregate(want, want.unit_seg_off)
    assert np.array_equal(gc, wc)
    assert_f32_bits_equal(gmn, wmn, name + " min")
    assert_f32_bits_equal(gmx, wmx, name + " max")
    assert np.array_equal(gsm.view(np.uint64), wsm.view(np.uint64))  # same fold order on the host


@pytest.mark.parametrize("chunk_len", [8, 64, 256, 1000, 4096])
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_chunk_speculative_chain_is_the_sequential_chain(oracle, case, chunk_len):
    """The parallel (chunked, speculative, fixpoint) segmentation must give exactly the rows of the
    sequential chain, for any chunk length -- including chunks far shorter than a segment."""
    name, ts, vals, off, ebs = case
    want = oracle.compress(ts, vals, off, eb=ebs)
    got = emu.compress(ts, vals, off, eb=ebs, chunk_len=chunk_len)
    assert_segments_equal(got, want, f"{name} chunk_len={chunk_len}")


@pytest.mark.parametrize("chunk_len", [16, 128, 1024])
def test_chunk_speculation_with_models_longer_than_chunks(oracle, chunk_len):
    """Constant / linear runs much longer than a chunk: speculative chains hit their budget, are cut
    short and resumed; models span (skip) whole chunks."""
    rng = np.random.default_rng(3)
    parts = [np.full(5000, 3.25, np.float32), rng.uniform(0, 1, 37).astype(np.float32), np.full(3000, -7.5, np.float32),
             (0.5 * np.arange(4000)).astype(np.float32), rng.uniform(0, 1, 300).astype(np.float32), np.full(2500, 1.0, np.float32)]
    vals = np.concatenate(parts)
    ts = syn.regular_timestamps(len(vals))
    for eb in ((0, 0.0), (2, 1.0), (1, 0.1)):
        rounds = []
        want = oracle.compress(ts, vals, eb=eb)
        got = emu.compress(ts, vals, eb=eb, chunk_len=chunk_len, rounds=rounds)
        assert_segments_equal(got, want, f"long models eb={eb} chunk_len={chunk_len}")
        assert rounds[0] >= 2  # speculation really was exercised


def test_chunk_speculation_converges_quickly_on_benchmark_like_data(oracle):
    ts, vals, off = syn.multi_series(4, 60_000, 11, "sine")
    # 1 % and lossless: chains re-synchronise within a chunk -> speculation, one repair round, done.
    # 5 % on this very smooth signal: chains anchored at different points never meet, the fixpoint
    # degrades to (at worst) one chunk per round -- still exact, and bounded by the chunk count.
    for eb, max_rounds in (((2, 1.0), 3), ((0, 0.0), 1), ((2, 5.0), 8 + 2)):
        rounds = []
        want = oracle.compress(ts, vals, off, eb=eb)
        got = emu.compress(ts, vals, off, eb=eb, chunk_len=8192, rounds=rounds)
        assert_segments_equal(got, want, f"sine eb={eb}")
        assert rounds[0] <= max_rounds, (eb, rounds)


# ---- the asynchronous scheduler (sched_advance + the worker loop of k_spec_async), stepped in random orders ----------

@pytest.mark.parametrize("in_flight", [1, 3, 16])
@pytest.mark.parametrize("chunk_len", [8, 64, 1000])
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_async_scheduler_gives_the_sequential_chain(oracle, case, chunk_len, in_flight):
    """Whatever the order in which workers claim chunks and chains complete, the per-unit frontier must end on
    exactly the sequential chain (and never stall or overflow the queue)."""
    name, ts, vals, off, ebs = case
    want = oracle.compress(ts, vals, off, eb=ebs)
    for seed in (1, 2, 3):
        got = emu.compress(ts, vals, off, eb=ebs, chunk_len=chunk_len, sched_seed=seed, in_flight=in_flight)
        assert_segments_equal(got, want, f"{name} chunk_len={chunk_len} in_flight={in_flight} seed={seed}")


@pytest.mark.parametrize("chunk_len", [16, 128, 1024])
def test_async_scheduler_with_models_longer_than_chunks(oracle, chunk_len):
    rng = np.random.default_rng(3)
    parts = [np.full(5000, 3.25, np.float32), rng.uniform(0, 1, 37).astype(np.float32), np.full(3000, -7.5, np.float32),
             (0.5 * np.arange(4000)).astype(np.float32), rng.uniform(0, 1, 300).astype(np.float32), np.full(2500, 1.0, np.float32)]
    vals = np.concatenate(parts)
    ts = syn.regular_timestamps(len(vals))
    for eb in ((0, 0.0), (2, 1.0), (1, 0.1)):
        want = oracle.compress(ts, vals, eb=eb)
        for seed, in_flight in ((1, 1), (2, 4), (3, 64), (4, 1000)):
            got = emu.compress(ts, vals, eb=eb, chunk_len=chunk_len, sched_seed=seed, in_flight=in_flight)
            assert_segments_equal(got, want, f"long models eb={eb} chunk_len={chunk_len} seed={seed}")


def test_async_scheduler_runs_each_chunk_once_when_workers_are_scarce(oracle):
    """One worker: every later chunk is still unstarted when the frontier reaches it and is re-aimed at its exact
    entry, so no chain is ever run speculatively and discarded (chain runs == chunks with a fit start)."""
    ts, vals, off = syn.multi_series(2, 40_000, 5, "sine")
    runs = []
    want = oracle.compress(ts, vals, off, eb=(2, 1.0))
    got = emu.compress(ts, vals, off, eb=(2, 1.0), chunk_len=4096, rounds=runs, sched_seed=9, in_flight=1)
    assert_segments_equal(got, want, "one worker")
    assert runs[0] <= 10, runs  # 40 000 / 4096 -> 10 chunks per unit


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_eight_point_screen_agrees_with_fit_next_model(case):
    """skip_rejected (warp engine) skips a start iff neither model can take its first 8 points; that must be exactly
    the set of starts whose fit the reference rejects (bytes per value > 4)."""
    name, ts, vals, off, ebs = case
    for u in range(len(off) - 1):
        a, b = int(off[u]), int(off[u + 1])
        if b - a == 0:
            continue
        assert emu.check_eight_points(ts[a:b], vals[a:b], ebs[u]) == 0, (name, u)


@pytest.mark.parametrize("eb", [(0, 0.0), (1, 0.05)], ids=["lossless", "abs0.05"])
def test_macaque_v_word_reader_at_every_alignment(oracle, eb):
    """The MacaqueV decoder reads its stream as aligned 16-byte quads (WordBitReader, mdb_device.cuh); the bytes before
    and after the stream in those quads must neither be read as data nor matter.  Streams of 1 to ~300 bytes (values and
    residuals), shifted to each of the 16 positions inside a quad, with poisoned bytes around them."""
    rng = np.random.default_rng(5)
    units = [rng.uniform(-50, 50, n).astype(np.float32) for n in (1, 2, 3, 5, 8, 13, 40, 77)]
    units.append(np.concatenate([np.full(30, 2.5, np.float32), rng.uniform(-50, 50, 9).astype(np.float32)]))  # a model, then residuals
    vals = np.concatenate(units)
    ts = np.concatenate([syn.regular_timestamps(len(u)) for u in units])
    off = np.concatenate([[0], np.cumsum([len(u) for u in units])]).astype(np.uint64)
    want = oracle.compress(ts, vals, off, eb=eb)
    wts, wval, _ = oracle.grid(want)
    wsum = oracle.segment_sums(want)
    for shift in range(16):
        cols = {c: getattr(want, c) for c in oracle._COLS}
        for name in ("values", "residuals"):
            data = cols[name + "_data"]
            backing = np.full(len(data) + 64, 0xA5, np.uint8)  # 64-byte aligned start is not guaranteed: align by hand
            start = (-backing.ctypes.data) % 16 + shift
            backing[start:start + len(data)] = data
            cols[name + "_data"] = backing[start:start + len(data)]
            assert cols[name + "_data"].ctypes.data % 16 == shift % 16
        shifted = oracle.Segments(**cols)
        gts, gval, _ = emu.grid(shifted)
        assert np.array_equal(gts, wts), shift
        assert gval.tobytes() == wval.tobytes(), shift
        assert emu.segment_sums(shifted)[0].tobytes() == wsum.tobytes(), shift


@pytest.mark.parametrize("block", range(4))
def test_grid_and_sum_bodies_fuzz(oracle, block):
    """Random stitched series (constants, ramps, noise, repeats, signed zeros, NaN / inf), random bounds, regular or
    irregular timestamps: the grid and sum bodies reproduce the oracle from segments whose byte columns sit at a
    random alignment between poisoned bytes."""
    from tests.test_warp_fit_emulated import _fuzz_series
    for seed in range(40 * block, 40 * block + 40):
        rng = np.random.default_rng(77000 + seed)
        vals = _fuzz_series(rng)
        n = len(vals)
        step = rng.integers(1, 2000, n) if seed % 3 == 0 else np.full(n, int(rng.integers(1, 5000)))
        ts = (int(rng.integers(0, 2_000_000_000_000_000)) + np.cumsum(step)).astype(np.int64)
        eb = [(0, 0.0), (1, float(10.0 ** rng.integers(-3, 3))), (2, float(rng.choice([0.01, 0.5, 1.0, 5.0, 30.0, 100.0])))][seed % 3]
        want = oracle.compress(ts, vals, eb=eb)
        wts, wval, _ = oracle.grid(want)
        cols = {c: getattr(want, c) for c in oracle._COLS}
        for name in ("values", "residuals", "timestamps"):
            data = cols[name + "_data"]
            backing = np.full(len(data) + 64, 0x5A, np.uint8)
            start = (-backing.ctypes.data) % 16 + int(rng.integers(0, 16))
            backing[start:start + len(data)] = data
            cols[name + "_data"] = backing[start:start + len(data)]
        seg = oracle.Segments(**cols)
        gts, gval, _ = emu.grid(seg)
        assert np.array_equal(gts, wts), seed
        assert_f32_bits_equal(gval, wval, f"fuzz grid seed={seed}")
        assert_f32_bits_equal(emu.segment_sums(seg)[0], oracle.segment_sums(want), f"fuzz sums seed={seed}", nan_payload_matters=False)
