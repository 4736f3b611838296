"""`rewrite_least_mantissa_bits` (macaque_v.rs:168-196) picks the mantissa bits to clear from
`23 - factorized_epsilon.log2().abs().floor() as i32` (macaque_v.rs:185).  f32::log2 is the platform libm's log2f; the
CUDA library computes log2 in f64 and rounds once to f32 (mdb_device.cuh: rewrite_position).  The position is a step
function of the input, so two forms that change at the same bit patterns to the same values are the same function: this
compares the two forms on EVERY non-negative f32 pattern from +0 to +inf (2^31 - 2^23 + 1 inputs, in C++ on all cores).
The device's own log2 is compared the same way in tests/test_gpu_parity.py::test_rewrite_position_on_the_device."""
import numpy as np


def test_f64_log2_gives_the_reference_position_for_every_f32(oracle):
    bits_ref, pos_ref = oracle.rewrite_position_steps(0)
    bits_f64, pos_f64 = oracle.rewrite_position_steps(1)
    assert len(bits_ref) > 250  # one step per binade, roughly
    assert np.array_equal(bits_ref, bits_f64)
    assert np.array_equal(pos_ref, pos_f64)


def test_steps_agree_with_the_pointwise_functions(oracle):
    bits, pos = oracle.rewrite_position_steps(0, 0x3F000000, 0x40000000, n_threads=3)
    L = oracle.lib()
    rng = np.random.default_rng(0)
    probe = np.concatenate([bits, bits - 1, bits + 1, rng.integers(0x3F000000, 0x40000001, 2000).astype(np.uint32)])
    probe = probe[(probe >= 0x3F000000) & (probe <= 0x40000000)]
    for b in probe:
        want = L.mdbo_rewrite_position_libm(float(np.array([b], np.uint32).view(np.float32)[0]))
        k = np.searchsorted(bits, b, side="right") - 1
        assert pos[k] == want, hex(int(b))


def test_nan_and_negative_inputs(oracle):
    """Not reachable from rewrite_least_mantissa_bits (the bound is non-negative and finite), but both forms agree."""
    L = oracle.lib()
    for b in (0x7FC00000, 0x7F800001, 0xFFC00000, 0x80000000, 0xBF800000, 0xFF800000):
        x = float(np.array([b], np.uint32).view(np.float32)[0])
        assert L.mdbo_rewrite_position_libm(x) == L.mdbo_rewrite_position_f64(x)
