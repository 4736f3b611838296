"""Pins the CPU oracle against every golden vector / known-answer test the reference holds for the
hot path.  Each test cites the reference test it restates; paths are relative to
/root/reference/crates/modelardb_compression/src/.  Nothing here reads /root/reference at run time.
"""
import math

import numpy as np
import pytest
from hypothesis import given, settings
from hypothesis import strategies as st

from modelardb_rs_b200 import synthetic

LOSSLESS = (0, 0.0)
F32_MAX = float(np.finfo(np.float32).max)
F32_MIN = -F32_MAX  # Rust's f32::MIN
ABS_MAX = (1, F32_MAX)      # modelardb_test/src/lib.rs:49
REL_MAX = (2, 100.0)        # modelardb_test/src/lib.rs:52
ABS5, REL5 = (1, 5.0), (2, 5.0)
ABS10, REL10 = (1, 10.0), (2, 10.0)

any_f32 = st.floats(width=32, allow_nan=True, allow_infinity=True)
INF, NAN = float("inf"), float("nan")


def eq_or_nan(a, b):
    return a == b or (math.isnan(a) and math.isnan(b))


# ------------------------------------------------------------------ models/bits.rs:188-299
def test_bit_order_and_padding(oracle):
    # bits.rs: MSB-first; finish() pads with zeros, finish_with_one_bits() with ones.  Observed through
    # the two public encoders: a lossless MacaqueV stream of [37, 73] and an irregular timestamp stream.
    r = oracle.macaque_v_compress(LOSSLESS, [73.0, 37.0, 37.0, 37.0, 73.0])
    # compression.rs:933-978 (KAT) expects 8 bytes; the byte values were derived by hand from the format:
    # 73.0 raw = 0x42920000, then `11`+lz(8)+len(7)+bits, `10`, `10`, `0`+7 bits, zero padded.
    assert list(r.data) == [66, 146, 0, 0, 208, 60, 58, 67]


# ------------------------------------------------------------------ models/mod.rs:298-476
@settings(max_examples=200, deadline=None)
@given(any_f32)
def test_same_value_is_always_within_lossless_error_bound(oracle, v):  # mod.rs:300-303
    assert oracle.is_value_within_error_bound(LOSSLESS, v, v)


@settings(max_examples=200, deadline=None)
@given(any_f32)
def test_non_finite_values_are_never_within_max_bounds_of_other_values(oracle, v):  # mod.rs:305-387
    for special in (INF, -INF, NAN):
        if eq_or_nan(v, special):
            continue
        for eb in (ABS_MAX, REL_MAX):
            assert not oracle.is_value_within_error_bound(eb, special, v)
            assert not oracle.is_value_within_error_bound(eb, v, special)


def test_different_value_is_within_non_zero_error_bounds(oracle):  # mod.rs:389-405
    assert oracle.is_value_within_error_bound((1, 1.0), 10.0, 11.0)
    assert oracle.is_value_within_error_bound(REL10, 10.0, 11.0)


def test_len(oracle):  # mod.rs:408-416
    assert oracle.length(1658671178037, 1658671178037, b"") == 1
    assert oracle.length(1658671178037, 1658671187047, bytes([10])) == 10


def test_split_into_models_and_residuals(oracle):  # mod.rs:432-464, :467-475
    ts = oracle.decompress_all_timestamps(100, 500, bytes([5]))
    assert list(ts) == [100, 200, 300, 400, 500]
    # residuals [.., 2] -> the last two timestamps belong to the residuals
    seg = oracle.Segments(
        model_type_id=np.array([0], np.int8), start_time=np.array([100], np.int64), end_time=np.array([500], np.int64),
        min_value=np.array([10.0], np.float32), max_value=np.array([10.0], np.float32),
        timestamps_off=np.array([0, 1], np.uint64), timestamps_data=np.array([5], np.uint8),
        values_off=np.array([0, 0], np.uint64), values_data=np.zeros(0, np.uint8),
        residuals_off=np.array([0, 0], np.uint64), residuals_data=np.zeros(0, np.uint8))
    ts, val, off = oracle.grid(seg)
    assert list(ts) == [100, 200, 300, 400, 500] and list(val) == [10.0] * 5 and list(off) == [0, 5]


# ------------------------------------------------------------------ models/timestamps.rs:303-405
def _roundtrip_ts(oracle, ts, expected_len=None):
    b = oracle.compress_residual_timestamps(ts)
    if expected_len is not None:
        assert len(b) == expected_len
    out = oracle.decompress_all_timestamps(ts[0], ts[-1], b)
    assert list(out) == list(ts)
    assert oracle.length(ts[0], ts[-1], b) == len(ts)
    return b


def test_compress_timestamps_zero_one_two(oracle):  # timestamps.rs:303-318
    assert oracle.compress_residual_timestamps([]) == b""
    assert oracle.compress_residual_timestamps([100]) == b""
    assert oracle.compress_residual_timestamps([100, 300]) == b""


def test_compress_and_decompress_known_timestamps(oracle):  # timestamps.rs:320-404
    _roundtrip_ts(oracle, [1579701905500, 1579701905600, 1579701905700, 1579701905800, 1579701905900], 1)
    _roundtrip_ts(oracle, [1579694400057, 1579694400197, 1579694400353, 1579694400493, 1579694400650], 4)
    assert _roundtrip_ts(oracle, [100, 100, 200], 1) == bytes([0b10111111])
    _roundtrip_ts(oracle, [100, 37, 38, 200], 3)
    _roundtrip_ts(oracle, [500, 245, 246, 500], 4)
    _roundtrip_ts(oracle, [5000, 2953, 2954, 5000], 5)
    _roundtrip_ts(oracle, [5000000000, 2852516353, 2852516354, 5000000000], 10)


def test_regular_length_encoding(oracle):  # timestamps.rs:99-108, SURVEY appendix A
    for n, expect in ((5, [5]), (127, [127]), (128, [0, 128]), (65536, [1, 0, 0])):
        assert list(oracle.compress_residual_timestamps(synthetic.regular_timestamps(n))) == expect


def test_generated_timestamps_roundtrip(oracle):  # timestamps.rs:406-416
    _roundtrip_ts(oracle, synthetic.regular_timestamps(1000, 0, 100))
    _roundtrip_ts(oracle, synthetic.irregular_timestamps(1000, 7, 0))


@settings(max_examples=100, deadline=None)
@given(st.lists(st.integers(min_value=-(2**62), max_value=2**62), min_size=3, max_size=60, unique=True))
def test_sorted_random_timestamps_roundtrip(oracle, ts):  # timestamps.rs:418-430 (proptest)
    _roundtrip_ts(oracle, sorted(ts))


# ------------------------------------------------------------------ models/pmc_mean.rs:119-335
@settings(max_examples=200, deadline=None)
@given(any_f32)
def test_pmc_can_fit_sequence_of_value_lossless(oracle, v):  # pmc_mean.rs:119-152
    n, mean = oracle.pmc_fit_prefix(LOSSLESS, [v] * 5)
    assert n == 5 and eq_or_nan(float(mean), float(np.float32(v)))


def test_pmc_specials(oracle):
    for v in (INF, -INF, NAN):
        n, mean = oracle.pmc_fit_prefix(LOSSLESS, [v] * 5)
        assert n == 5 and eq_or_nan(float(mean), v)


@settings(max_examples=200, deadline=None)
@given(any_f32)
def test_pmc_cannot_mix_non_finite_and_other_values(oracle, v):  # pmc_mean.rs:160-262
    for special in (INF, -INF, NAN):
        if eq_or_nan(v, special):
            continue
        for eb in (ABS_MAX, REL_MAX):
            assert oracle.pmc_fit_prefix(eb, [v, special])[0] == 1
            assert oracle.pmc_fit_prefix(eb, [special, v])[0] == 1


def test_pmc_fit_matrix(oracle):  # pmc_mean.rs:265-299
    seq = [42.0, 42.0, 42.8, 42.0, 42.0]
    assert oracle.pmc_fit_prefix(LOSSLESS, seq)[0] < 5
    assert oracle.pmc_fit_prefix(ABS5, seq)[0] == 5
    assert oracle.pmc_fit_prefix(REL5, seq)[0] == 5


# ------------------------------------------------------------------ models/swing.rs:366-798
START_TIME, SAMPLING = 1658671178037, 1000


def _ts(n):
    return [START_TIME + i * SAMPLING for i in range(n)]


@settings(max_examples=200, deadline=None)
@given(any_f32)
def test_swing_can_fit_sequence_of_value_lossless(oracle, v):  # swing.rs:366-412
    n, first, last = oracle.swing_fit_prefix(LOSSLESS, _ts(5), [v] * 5)
    assert n == 5
    v32 = float(np.float32(v))
    if math.isnan(v32):
        assert math.isnan(float(first)) and math.isnan(float(last))
    else:
        assert float(first) == v32 and float(last) == v32


@settings(max_examples=100, deadline=None)
@given(st.floats(width=32, allow_nan=False, allow_infinity=False, allow_subnormal=False),
       st.floats(width=32, allow_nan=False, allow_infinity=False, allow_subnormal=False))
def test_swing_can_fit_two_finite_values(oracle, a, b):  # swing.rs:420-428
    assert oracle.swing_fit_prefix(LOSSLESS, _ts(2), [a, b])[0] == 2


@settings(max_examples=200, deadline=None)
@given(any_f32)
def test_swing_cannot_mix_non_finite_and_other_values(oracle, v):  # swing.rs:430-537
    for special in (INF, -INF, NAN):
        if eq_or_nan(v, special):
            continue
        for eb in (ABS_MAX, REL_MAX):
            assert oracle.swing_fit_prefix(eb, _ts(2), [v, special])[0] == 1
            assert oracle.swing_fit_prefix(eb, _ts(2), [special, v])[0] == 1


def test_swing_fit_matrix(oracle):  # swing.rs:539-570
    assert oracle.swing_fit_prefix(LOSSLESS, _ts(5), [42.0, 84.0, 126.0, 168.0, 210.0])[0] == 5
    seq = [42.0, 42.0, 42.8, 42.0, 42.0]
    assert oracle.swing_fit_prefix(LOSSLESS, _ts(5), seq)[0] < 5
    assert oracle.swing_fit_prefix(ABS5, _ts(5), seq)[0] == 5
    assert oracle.swing_fit_prefix(REL5, _ts(5), seq)[0] == 5


SWING_SEQ = [42.0, 42.0, 42.8, 42.0, 41.0, 40.0, 42.0, 42.0, 42.0, 42.1]


def _slope_icpt(t0, v0, t1, v1):  # swing.rs:323-340
    if v0 == v1:
        return 0.0, v0
    s = (v1 - v0) / float(t1 - t0)
    return s, v0 - s * float(t0)


def test_swing_slope_is_between_hyperplanes(oracle):  # swing.rs:572-577, :647-666
    ts = _ts(len(SWING_SEQ))
    n, first, last = oracle.swing_fit_prefix(REL5, ts, SWING_SEQ)
    assert n == len(SWING_SEQ)
    ls, _, us, _ = oracle.swing_bounds(REL5, ts, SWING_SEQ)
    end_time = START_TIME + len(SWING_SEQ) * SAMPLING
    slope, _ = _slope_icpt(START_TIME, float(first), end_time, float(last))
    assert ls <= slope <= us


def test_swing_can_minimize_mse(oracle):  # swing.rs:580-585, :600-645
    ts = _ts(len(SWING_SEQ))
    n, first, last = oracle.swing_fit_prefix(REL5, ts, SWING_SEQ)
    assert n == len(SWING_SEQ)
    end_time = START_TIME + len(SWING_SEQ) * SAMPLING
    ds, di = _slope_icpt(START_TIME, float(np.float32(SWING_SEQ[0])), end_time, float(np.float32(SWING_SEQ[-1])))
    ms, mi = _slope_icpt(START_TIME, float(first), end_time, float(last))
    mse = sum((float(np.float32(v)) - (ds * t + di)) ** 2 for v, t in zip(SWING_SEQ, ts))
    opt = sum((float(np.float32(v)) - (ms * t + mi)) ** 2 for v, t in zip(SWING_SEQ, ts))
    assert mse > opt


def _one_row(oracle, **kw):
    d = dict(model_type_id=1, start_time=0, end_time=0, timestamps=b"", min_value=0.0, max_value=0.0, values=b"",
             residuals=b"")
    d.update(kw)
    u8 = lambda b: np.frombuffer(b, np.uint8).copy()
    return oracle.Segments(
        model_type_id=np.array([d["model_type_id"]], np.int8), start_time=np.array([d["start_time"]], np.int64),
        end_time=np.array([d["end_time"]], np.int64), min_value=np.array([d["min_value"]], np.float32),
        max_value=np.array([d["max_value"]], np.float32),
        timestamps_off=np.array([0, len(d["timestamps"])], np.uint64), timestamps_data=u8(d["timestamps"]),
        values_off=np.array([0, len(d["values"])], np.uint64), values_data=u8(d["values"]),
        residuals_off=np.array([0, len(d["residuals"])], np.uint64), residuals_data=u8(d["residuals"]))


@settings(max_examples=100, deadline=None)
@given(st.integers(-2**31, 2**31 - 1), st.integers(-2**31, 2**31 - 1))
def test_swing_sum(oracle, a, b):  # swing.rs:668-677: sum(START, END, [], first, last, 0) == first + last
    f = np.float32(int(math.fmod(a, 1_000_000)))
    l = np.float32(int(math.fmod(b, 1_000_000)))
    mn, mx = min(f, l), max(f, l)
    values = b"" if f <= l and not (f == l) else (b"" if f < l else b"\x00")
    if f == l:
        values = b""
    seg = _one_row(oracle, start_time=START_TIME, end_time=START_TIME + SAMPLING, min_value=mn, max_value=mx,
                   values=values)
    assert oracle.segment_sums(seg)[0] == np.float32(f + l)


@settings(max_examples=100, deadline=None)
@given(st.integers(-2**31, 2**31 - 1))
def test_swing_grid_constant(oracle, a):  # swing.rs:680-708
    v = np.float32(int(math.fmod(a, 1_000_000)))
    seg = _one_row(oracle, start_time=START_TIME, end_time=START_TIME + SAMPLING, min_value=v, max_value=v)
    ts, val, _ = oracle.grid(seg)
    assert list(ts) == [START_TIME, START_TIME + SAMPLING] and list(val) == [v, v]


@pytest.mark.parametrize("reverse", [False, True])
def test_swing_reconstructs_linear_sequence_exactly(oracle, reverse):  # swing.rs:717-798
    values = [float(v) for v in range(42, 4201, 42)]
    if reverse:
        values.reverse()
    ts = _ts(len(values))
    seg = oracle.compress(ts, values, eb=LOSSLESS)
    assert len(seg) == 1 and seg.model_type_id[0] == oracle.SWING
    gts, gval, _ = oracle.grid(seg)
    assert list(gts) == ts
    assert list(gval) == values


# ------------------------------------------------------------------ models/macaque_v.rs:348-545
def test_macaque_v_empty(oracle):  # macaque_v.rs:348-351
    assert oracle.macaque_v_compress(LOSSLESS, []).data == b""


@settings(max_examples=200, deadline=None)
@given(any_f32)
def test_macaque_v_single_and_repeated_value(oracle, v):  # macaque_v.rs:353-376
    for seq in ([v], [v, v]):
        r = oracle.macaque_v_compress(LOSSLESS, seq)
        assert eq_or_nan(float(r.last_value), float(np.float32(v)))
        assert (r.last_leading_zero_bits, r.last_trailing_zero_bits) == (255, 0)


def test_macaque_v_window_state(oracle):  # macaque_v.rs:378-398
    for seq in ([37.0, 73.0], [37.0, 71.0, 73.0]):
        r = oracle.macaque_v_compress(LOSSLESS, seq)
        assert float(r.last_value) == 73.0
        assert (r.last_leading_zero_bits, r.last_trailing_zero_bits) == (8, 17)


@pytest.mark.parametrize("eb", [ABS10, REL10])
def test_macaque_v_value_within_bound_reuses_previous(oracle, eb):  # macaque_v.rs:401-433
    before = oracle.macaque_v_compress(eb, [10.0])
    after = oracle.macaque_v_compress(eb, [10.0, 11.0])
    assert float(after.last_value) == float(before.last_value) == 10.0
    assert (after.last_leading_zero_bits, after.last_trailing_zero_bits) == (255, 0)


bits32 = st.integers(0, 2**32 - 1).map(lambda b: np.array([b], np.uint32).view(np.float32)[0])


@settings(max_examples=200, deadline=None)
@given(st.lists(bits32, min_size=1, max_size=49), st.booleans())
def test_macaque_v_lossless_roundtrip(oracle, values, seeded):  # macaque_v.rs:436-521 (sum and grid proptests)
    vals = np.array(values, np.float32)
    seed = np.float32(37.0) if seeded else None
    r = oracle.macaque_v_compress(LOSSLESS, vals, seed=seed)
    out = oracle.macaque_v_grid(r.data, len(vals), seed=seed)
    assert out.view(np.uint32).tolist() == vals.view(np.uint32).tolist()
    expected = np.float32(0.0) if seeded else vals[0]
    with np.errstate(all="ignore"):
        for v in (vals if seeded else vals[1:]):
            expected = np.float32(expected + v)
    got = oracle.macaque_v_sum(r.data, len(vals), seed=seed)
    assert eq_or_nan(float(got), float(expected))


def test_macaque_v_sum_single_values(oracle):  # macaque_v.rs:449-464
    r = oracle.macaque_v_compress(LOSSLESS, [37.0])
    assert oracle.macaque_v_sum(r.data, 1) == 37.0
    r = oracle.macaque_v_compress(LOSSLESS, [37.0], seed=37.0)
    assert oracle.macaque_v_sum(r.data, 1, seed=37.0) == 37.0


# ------------------------------------------------------------------ types.rs:535-890
TS5 = [100, 200, 300, 400, 500]


def _fit_and_finish(oracle, values, model_type, end_index, model_min, model_max, model_values_len, seg_min, seg_max,
                    seg_values_len):  # types.rs:791-860
    m = oracle.fit_next_model(0, LOSSLESS, TS5, values)
    assert m.model_type_id == model_type and m.start_index == 0 and m.end_index == end_index
    assert float(m.min_value) == model_min and float(m.max_value) == model_max
    assert len(m.values) == model_values_len
    seg = oracle.model_finish(m, LOSSLESS, len(TS5) - 1, TS5, values)
    assert len(seg) == 1
    row = seg.row(0)
    assert float(row["min_value"]) == float(np.float32(seg_min))
    assert float(row["max_value"]) == float(np.float32(seg_max))
    assert len(row["values"]) == seg_values_len
    return m, row


@pytest.mark.parametrize("values,end,seg_min,seg_max,seg_len", [
    ([10.0, 10.0, 10.0, 10.0, 10.0], 4, 10.0, 10.0, 0),        # types.rs:535-547
    ([10.0, 10.0, 10.0, 10.0, F32_MIN], 3, F32_MIN, 10.0, 1),  # types.rs:549-561
    ([10.0, 10.0, 10.0, 10.0, F32_MAX], 3, 10.0, F32_MAX, 0),  # types.rs:563-575
    ([10.0, 10.0, 10.0, F32_MIN, F32_MAX], 2, F32_MIN, F32_MAX, 4),  # types.rs:577-589
])
def test_encoding_decoding_for_pmc_mean(oracle, values, end, seg_min, seg_max, seg_len):
    _, row = _fit_and_finish(oracle, values, oracle.PMC_MEAN, end, 10.0, 10.0, 0, seg_min, seg_max, seg_len)
    assert oracle.decode_values_for_pmc_mean(row["min_value"], row["max_value"], row["values"]) == 10.0


@pytest.mark.parametrize("values,end,mmin,mmax,mlen,seg_min,seg_max,seg_len", [
    ([10.0, 20.0, 30.0, 40.0, 50.0], 4, 10.0, 50.0, 0, 10.0, 50.0, 0),            # types.rs:627-640
    ([10.0, 20.0, 30.0, 40.0, F32_MIN], 3, 10.0, 40.0, 0, F32_MIN, 40.0, 5),      # types.rs:642-655
    ([10.0, 20.0, 30.0, 40.0, F32_MAX], 3, 10.0, 40.0, 0, 10.0, F32_MAX, 5),      # types.rs:657-670
    ([10.0, 20.0, 30.0, F32_MIN, F32_MAX], 2, 10.0, 30.0, 0, F32_MIN, F32_MAX, 8),  # types.rs:672-685
    ([50.0, 40.0, 30.0, 20.0, 10.0], 4, 10.0, 50.0, 1, 10.0, 50.0, 1),            # types.rs:687-700
    ([50.0, 40.0, 30.0, 20.0, F32_MIN], 3, 20.0, 50.0, 1, F32_MIN, 50.0, 5),      # types.rs:702-715
    ([50.0, 40.0, 30.0, 20.0, F32_MAX], 3, 20.0, 50.0, 1, 20.0, F32_MAX, 5),      # types.rs:717-730
    ([50.0, 40.0, 30.0, F32_MIN, F32_MAX], 2, 30.0, 50.0, 1, F32_MIN, F32_MAX, 8),  # types.rs:732-745
])
def test_encoding_decoding_for_swing(oracle, values, end, mmin, mmax, mlen, seg_min, seg_max, seg_len):
    m, row = _fit_and_finish(oracle, values, oracle.SWING, end, mmin, mmax, mlen, seg_min, seg_max, seg_len)
    first, last = oracle.decode_values_for_swing(row["min_value"], row["max_value"], row["values"])
    assert float(first) == values[m.start_index] and float(last) == values[m.end_index]  # types.rs:776-787


def test_model_with_fewest_bytes_is_selected(oracle):  # types.rs:862-890
    ts = synthetic.regular_timestamps(50, 0, 100)
    values = np.concatenate([synthetic.constant(25, 3), synthetic.uniform_random(25, 4, 0.0, 100.0)])
    m = oracle.fit_next_model(0, REL10, ts, values)
    assert m.model_type_id == oracle.PMC_MEAN


# ------------------------------------------------------------------ compression.rs:421-978
def _assert_roundtrip_within_bound(oracle, eb, ts, values, seg):  # compression.rs:865-929
    gts, gval, _ = oracle.grid(seg)
    assert gts.tolist() == list(np.asarray(ts, np.int64))
    assert len(gval) == len(values)
    v32 = np.asarray(values, np.float32)
    for real, approx in zip(v32, gval):
        assert oracle.is_value_within_error_bound(eb, real, approx), (real, approx, eb)


def test_try_compress_empty(oracle):  # compression.rs:422-434
    assert len(oracle.compress([], [], eb=LOSSLESS)) == 0


HALF_RANGE = (F32_MIN / 2.0, F32_MAX / 2.0)  # data_generation.rs:86-88


@pytest.mark.parametrize("irregular", [False, True])
@pytest.mark.parametrize("shape,eb,expected", [
    ("constant", LOSSLESS, [0]),          # compression.rs:436-454
    ("almost_constant", ABS5, [0]),       # compression.rs:456-494
    ("almost_constant", REL5, [0]),
    ("linear", LOSSLESS, [1]),            # compression.rs:496-514
    ("almost_linear", ABS5, [1]),         # compression.rs:516-554
    ("almost_linear", REL5, [1]),
    ("random", LOSSLESS, [2]),            # compression.rs:556-574
])
def test_try_compress_known_segment(oracle, irregular, shape, eb, expected):  # compression.rs:576-603
    ts = synthetic.irregular_timestamps(10, 11, 0) if irregular else synthetic.regular_timestamps(10, 0, 100)
    if shape == "constant":
        values = synthetic.constant(10, 5)
    elif shape == "almost_constant":
        values = synthetic.uniform_random(10, 5, 9.8, 10.2)
    elif shape == "linear":
        values = synthetic.linear(ts, 5)
    elif shape == "almost_linear":
        values = synthetic.linear(ts, 5, noise=(1.0, 1.05))
    else:
        values = synthetic.uniform_random(10, 5, *HALF_RANGE)
    seg = oracle.compress(ts, values, eb=eb)
    assert seg.model_type_id.tolist() == expected
    _assert_roundtrip_within_bound(oracle, eb, ts, values, seg)


@pytest.mark.parametrize("irregular", [False, True])
@pytest.mark.parametrize("generate,expected", [
    ([2, 1, 0], [2, 1, 0]),  # compression.rs:605-624
    ([0, 1, 2], [0, 1]),     # compression.rs:626-645: trailing random run becomes residuals of the Swing segment
])
def test_try_compress_known_time_series(oracle, irregular, generate, expected):  # compression.rs:647-707
    n = 50
    ts = synthetic.irregular_timestamps(3 * n, 13, 0) if irregular else synthetic.regular_timestamps(3 * n, 0, 100)
    parts = []
    for k, model in enumerate(generate):
        sl = ts[k * n:(k + 1) * n]
        if model == 0:
            parts.append(synthetic.constant(n, 21))
        elif model == 1:
            parts.append(synthetic.linear(sl, 22))
        else:
            parts.append(synthetic.uniform_random(n, 23, *HALF_RANGE))
    values = np.concatenate(parts)
    seg = oracle.compress(ts, values, eb=LOSSLESS)
    assert seg.model_type_id.tolist() == expected
    _assert_roundtrip_within_bound(oracle, LOSSLESS, ts, values, seg)


@pytest.mark.parametrize("irregular", [False, True])
@pytest.mark.parametrize("noise", [None, (1.0, 1.05)])
@pytest.mark.parametrize("eb", [LOSSLESS, ABS5, REL5])
def test_try_compress_synthetic_time_series(oracle, eb, irregular, noise):  # compression.rs:732-863
    ts, values = synthetic.mixed_series(50_000, 31, irregular=irregular, noise=noise)
    seg = oracle.compress(ts, values, eb=eb)
    gts, gval, _ = oracle.grid(seg)
    assert np.array_equal(gts, ts)
    ok = [oracle.is_value_within_error_bound(eb, r, a) for r, a in zip(values[::7], gval[::7])]
    assert all(ok)
    if eb == LOSSLESS:
        assert np.array_equal(gval.view(np.uint32), values.view(np.uint32))


def test_compress_and_store_residuals_in_a_separate_segment(oracle):  # compression.rs:932-978
    seg = oracle.macaque_v_segment(LOSSLESS, 0, 4, TS5, [73.0, 37.0, 37.0, 37.0, 73.0])
    assert len(seg) == 1
    row = seg.row(0)
    assert row["model_type_id"] == oracle.MACAQUE_V
    assert (row["start_time"], row["end_time"]) == (100, 500)
    assert row["timestamps"] == bytes([5])
    assert float(row["min_value"]) == 37.0 and float(row["max_value"]) == 73.0
    assert len(row["values"]) == 8 and row["residuals"] == b""


# ------------------------------------------------------------------ aggregates
def test_grouped_aggregates_known_answers(oracle):
    # crates/modelardb_embedded/src/operations/data_folder.rs:1166-1234 with the fixture at :2828-2848,
    # after the sort by (tags, time) of compression.rs:111-141: COUNT/MIN/MAX(field_1) and
    # SUM/AVG(field_2) GROUP BY tag_1 -> tag_x: 3, 37, 39, 75, 25; tag_y: 3, 71, 73, 165, 55.
    ts = np.array([100, 200, 300, 100, 200, 300], np.int64)
    field_1 = np.array([37.0, 38.0, 39.0, 73.0, 72.0, 71.0], np.float32)
    field_2 = np.array([24.0, 25.0, 26.0, 56.0, 55.0, 54.0], np.float32)
    unit_off = np.array([0, 3, 6], np.uint64)
    seg1 = oracle.compress(ts, field_1, unit_off, eb=LOSSLESS)
    seg2 = oracle.compress(ts, field_2, unit_off, eb=LOSSLESS)
    count, mn, mx, _ = oracle.aggregate(seg1, seg1.unit_seg_off)
    count2, _, _, sm = oracle.aggregate(seg2, seg2.unit_seg_off)
    assert count.tolist() == [3, 3]
    assert mn.tolist() == [37.0, 71.0] and mx.tolist() == [39.0, 73.0]
    assert sm.tolist() == [75.0, 165.0]
    assert (sm / count2).tolist() == [25.0, 55.0]


@pytest.mark.parametrize("eb", [LOSSLESS, ABS5, REL5])
def test_aggregates_from_segments_equal_aggregates_from_data_points(oracle, eb):
    # crates/modelardb_server/tests/integration_test.rs:1128-1246: COUNT/MIN/MAX equal, SUM/AVG within
    # 0.001 % of the same aggregate over the reconstructed data points.
    ts, values = synthetic.mixed_series(5000, 41, noise=(1.0, 1.05))
    seg = oracle.compress(ts, values, eb=eb)
    count, mn, mx, sm = oracle.aggregate(seg)
    _, gval, _ = oracle.grid(seg)
    assert count[0] == len(gval) == 5000
    assert mn[0] == gval.min() and mx[0] == gval.max()
    ref = float(gval.astype(np.float64).sum())
    assert abs(sm[0] - ref) <= 1e-5 * abs(ref)
