"""The warp-cooperative MacaqueV decoder and encoder (csrc/mdb_macaque_warp.cuh: staged stream, branch-free code walk,
parallel payload extraction, XOR scan; ballot loops for stored values and windows, scanned bit offsets, OR-ed stage) run
on the host by the warp emulator against the oracle's serial coder.  The GPU tests cover the same code through the
C-ABI; this is the inner loop for changing it on a machine without a GPU."""
import numpy as np
import pytest

from tests import emu_lib as emu


def _streams(rng, n):
    kind = int(rng.integers(0, 6))
    if kind == 0:
        return rng.uniform(-1e3, 1e3, n).astype(np.float32)                       # a new window at almost every value
    if kind == 1:
        return (100.0 + np.cumsum(rng.standard_normal(n))).astype(np.float32)     # random walk: reuse and new windows mixed
    if kind == 2:
        return np.repeat(rng.uniform(-5, 5, n // 7 + 1).astype(np.float32), 7)[:n]  # runs of equal values: `10` codes
    if kind == 3:
        return rng.choice(np.array([0.0, -0.0, 1.0, np.nan, np.inf, -np.inf, 3.4e38, 1e-45], np.float32), n)
    if kind == 4:
        return (rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32)).view(np.float32)  # arbitrary bit patterns
    return np.full(n, np.float32(rng.normal()), np.float32)


def _library(runs):
    """The product's decoder (takes runs of `0` / `10` codes in one step), or the build with the plain code walk
    (MDB_MACAQUE_SPECULATE_RUNS=0)."""
    return None if runs else emu.variant("MDB_MACAQUE_SPECULATE_RUNS=0")


@pytest.mark.parametrize("runs", [False, True], ids=["walk", "runs"])
@pytest.mark.parametrize("block", range(6))
def test_warp_decoder_equals_serial_decoder(oracle, block, runs):
    library = _library(runs)
    rng = np.random.default_rng(4100 + block)
    for case in range(25):
        n = int(rng.choice([1, 2, 31, 32, 33, 64, 65, 500, 1700, 5000])) if case % 2 else int(rng.integers(1, 3000))
        vals = _streams(rng, n)
        eb = [(0, 0.0), (1, float(10.0 ** rng.integers(-3, 2))), (2, float(rng.choice([0.1, 1.0, 10.0])))][case % 3]
        seed = None if case % 4 else np.float32(rng.normal())  # residual streams are seeded with the model's last value
        stream = oracle.macaque_v_compress(eb, vals, seed=seed)
        want = oracle.macaque_v_grid(stream.data, n, seed=seed)
        got, last = emu.warp_macaque_decode(stream.data, n, seed=seed, misalign=int(rng.integers(0, 16)), library=library)
        assert got.tobytes() == want.tobytes(), (block, case, n, eb, seed)
        assert np.float32(last).tobytes() == want[-1:].tobytes(), (block, case)


@pytest.mark.parametrize("runs", [False, True], ids=["walk", "runs"])
def test_warp_decoder_refills_its_stage(oracle, runs):
    """A stream much longer than the 2 KiB stage, decoded from every alignment."""
    rng = np.random.default_rng(9)
    vals = rng.uniform(-1e6, 1e6, 20_000).astype(np.float32)
    stream = oracle.macaque_v_compress((0, 0.0), vals)
    assert len(stream.data) > 30 * 2048
    for misalign in (0, 1, 2, 3, 7, 13):
        got, _ = emu.warp_macaque_decode(stream.data, len(vals), misalign=misalign, library=_library(runs))
        assert got.tobytes() == vals.tobytes(), misalign


def test_run_decoder_on_streams_that_switch_between_kinds_of_code(oracle):
    """Stretches of noise (`0` codes), of repeated values (`10` codes) and of values of wildly different magnitude (`11`
    codes) in one stream: the run decoder changes between its two modes and must stay exact across the switches."""
    rng = np.random.default_rng(12)
    parts = []
    for _ in range(60):
        kind, m = int(rng.integers(0, 3)), int(rng.integers(5, 300))
        if kind == 0:
            parts.append(100.0 + rng.standard_normal(m))
        elif kind == 1:
            parts.append(np.full(m, rng.normal()))
        else:
            parts.append(rng.standard_normal(m) * 10.0 ** rng.integers(-20, 20, m))
    vals = np.concatenate(parts).astype(np.float32)
    for eb in [(0, 0.0), (2, 1.0)]:
        for seed in (None, np.float32(3.5)):
            stream = oracle.macaque_v_compress(eb, vals, seed=seed)
            want = oracle.macaque_v_grid(stream.data, len(vals), seed=seed)
            got, _ = emu.warp_macaque_decode(stream.data, len(vals), seed=seed, misalign=3, library=_library(True))
            assert got.tobytes() == want.tobytes(), (eb, seed)


@pytest.mark.parametrize("block", range(6))
def test_warp_encoder_equals_serial_encoder(oracle, block):
    rng = np.random.default_rng(5200 + block)
    for case in range(25):
        n = int(rng.choice([1, 2, 31, 32, 33, 64, 65, 500, 1700, 5000])) if case % 2 else int(rng.integers(1, 3000))
        vals = _streams(rng, n)
        eb = [(0, 0.0), (1, float(10.0 ** rng.integers(-3, 2))), (2, float(rng.choice([0.1, 1.0, 10.0, 100.0])))][case % 3]
        want = oracle.macaque_v_compress(eb, vals)
        data, mn, mx, counted = emu.warp_macaque_encode(vals, eb)
        assert data == want.data, (block, case, n, eb)
        assert counted == len(want.data), (block, case)
        # min / max of the STORED values (macaque_v.rs:199-204); NaN payloads are not part of the contract
        for got, ref in ((mn, want.min_value), (mx, want.max_value)):
            assert (np.isnan(got) and np.isnan(ref)) or np.float32(got).tobytes() == np.float32(ref).tobytes(), (block, case, got, ref)


def test_warp_encoder_drains_its_stage(oracle):
    """A stream much longer than the 2 KiB stage (several drains, a partial word carried over each time), lossless and
    lossy; decoding the emulated encoder's bytes with the emulated decoder closes the loop."""
    rng = np.random.default_rng(10)
    vals = rng.uniform(-1e6, 1e6, 12_000).astype(np.float32)
    for eb in [(0, 0.0), (2, 0.5)]:
        want = oracle.macaque_v_compress(eb, vals)
        data, _, _, counted = emu.warp_macaque_encode(vals, eb)
        assert data == want.data and counted == len(data) and len(data) > 10 * 2048
        got, _ = emu.warp_macaque_decode(data, len(vals), misalign=5)
        assert got.tobytes() == oracle.macaque_v_grid(want.data, len(vals)).tobytes()


def test_wide_runs_of_reuse_codes_are_taken_and_exact(oracle):
    """Lossless high-entropy streams are fixed width after a warm-up (every code is `0` + the same number of bits): the
    decoder proves 256 codes at a time from their flag bits alone and scans once per 256 values.  Streams where runs break
    (a value repeated now and then, a window that has to widen) fall back to the 32-value batches and back."""
    rng = np.random.default_rng(77)
    lib = emu.lib()
    lib.emu_wide_runs.restype = __import__("ctypes").c_uint64
    walk = (100.0 + np.cumsum(rng.standard_normal(30_000))).astype(np.float32)
    broken = walk.copy()
    broken[5000:5003] = broken[4999]            # `10` codes inside a run
    broken[12_000] = np.float32(-3.0e38)        # an XOR that needs a new window
    noise = rng.uniform(-1e6, 1e6, 9000).astype(np.float32)
    for name, vals, expect_wide in (("walk", walk, True), ("broken", broken, True), ("noise", noise, True), ("short", walk[:200], False)):
        stream = oracle.macaque_v_compress((0, 0.0), vals)
        for misalign in (0, 3):
            before = lib.emu_wide_runs()
            got, last = emu.warp_macaque_decode(stream.data, len(vals), misalign=misalign)
            assert got.tobytes() == vals.tobytes(), (name, misalign)
            assert np.float32(last).tobytes() == vals[-1:].tobytes()
            assert (lib.emu_wide_runs() > before) == expect_wide, name
    # a seeded (residual-style) stream and a lossy one (runs of `10` codes: never wide)
    stream = oracle.macaque_v_compress((0, 0.0), walk[:5000], seed=np.float32(1.5))
    got, _ = emu.warp_macaque_decode(stream.data, 5000, seed=np.float32(1.5))
    assert got.tobytes() == walk[:5000].tobytes()
    lossy = oracle.macaque_v_compress((2, 1.0), walk)
    want = oracle.macaque_v_grid(lossy.data, len(walk))
    got, _ = emu.warp_macaque_decode(lossy.data, len(walk))
    assert got.tobytes() == want.tobytes()
