// mdb_oracle.cc -- CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY
// (see mdb_oracle.h).  Every function cites the reference file:line it follows; paths are
// relative to /root/reference/crates/modelardb_compression/src/ unless prefixed.
//
// Build: g++ -O2 -ffp-contract=off -fno-fast-math (Rust never contracts a*b+c into an FMA and
// x86-64 SSE2 has no excess precision, so plain float/double arithmetic is the same arithmetic).
#include "mdb_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

namespace {

using std::size_t;

// ---------------------------------------------------------------------------------------------
// Rust float semantics
// ---------------------------------------------------------------------------------------------

// f32::min / f64::min lower to llvm.minnum, which x86-64 LLVM selects as
// `isnan(a) ? b : (b < a ? b : a)` (X86ISelLowering combineFMinNumFMaxNum): NaN-ignoring, and on
// equal operands (+0.0 vs -0.0) the receiver `a` is returned.  Signed-zero ties are UNPINNED by any
// reference test; the GPU path uses the identical expression.
template <typename T> inline T rust_min(T a, T b) { return std::isnan(a) ? b : (b < a ? b : a); }
template <typename T> inline T rust_max(T a, T b) { return std::isnan(a) ? b : (b > a ? b : a); }

// `x as i32` for f32: saturating, NaN -> 0.
inline int32_t f32_as_i32(float x) {
    if (std::isnan(x)) return 0;
    if (x >= 2147483648.0f) return INT32_MAX;
    if (x <= -2147483648.0f) return INT32_MIN;
    return (int32_t)x;
}

inline uint32_t f32_bits(float v) { uint32_t u; std::memcpy(&u, &v, 4); return u; }
inline float f32_from_bits(uint32_t u) { float v; std::memcpy(&v, &u, 4); return v; }

struct ErrorBound { int kind; float value; };

// models/mod.rs:92-95
inline bool equal_or_nan(double v1, double v2) { return v1 == v2 || (std::isnan(v1) && std::isnan(v2)); }

// models/mod.rs:53-80
inline bool is_value_within_error_bound(ErrorBound eb, float real_value, float approximate_value) {
    if (equal_or_nan((double)real_value, (double)approximate_value)) return true;
    switch (eb.kind) {
    case MDBO_ABSOLUTE:
        return std::fabs(real_value - approximate_value) <= eb.value;
    case MDBO_RELATIVE: {
        float difference = real_value - approximate_value;
        float result = std::fabs(difference / real_value);
        return (result * 100.0f) <= eb.value;
    }
    default:
        return false;
    }
}

// models/mod.rs:83-90
inline double maximum_allowed_deviation(ErrorBound eb, double value) {
    switch (eb.kind) {
    case MDBO_ABSOLUTE: return (double)eb.value * 0.99;
    case MDBO_RELATIVE: return std::fabs(value * ((double)eb.value / 100.1));
    default: return 0.0;
    }
}

// ---------------------------------------------------------------------------------------------
// models/bits.rs
// ---------------------------------------------------------------------------------------------

// bits.rs:25-83 -- kept bit-at-a-time on purpose: this is the reference's reader and the timed CPU
// baseline must pay what the reference pays.
struct BitReader {
    size_t next_bit = 0;
    const uint8_t *bytes;
    size_t n_bytes;
    BitReader(const uint8_t *b, size_t n) : bytes(b), n_bytes(n) {}
    bool is_empty() const { return (next_bit / 8) == n_bytes; }           // bits.rs:45-47
    size_t remaining_bits() const { return 8 * n_bytes - next_bit; }      // bits.rs:50-52
    bool read_bit() { return read_bits(1) == 1; }                         // bits.rs:55-57
    uint64_t read_bits(uint8_t number_of_bits) {                          // bits.rs:61-82
        uint64_t value = 0;
        size_t start_bit = next_bit, end_bit = next_bit + number_of_bits;
        for (size_t bit = start_bit; bit < end_bit; bit++) {
            uint8_t byte = bytes[bit / 8];
            unsigned shift = 7 - (bit % 8);
            value = (value << 1) | ((uint64_t)(byte >> shift) & 1);
        }
        next_bit = end_bit;
        return value;
    }
};

// bits.rs:86-174
struct BitVecBuilder {
    uint8_t current_byte = 0;
    uint8_t remaining_bits = 8;
    std::vector<uint8_t> bytes;
    void append_a_zero_bit() { append_bits(0, 1); }
    void append_a_one_bit() { append_bits(1, 1); }
    void append_bits(uint64_t bits, uint8_t number_of_bits) {             // bits.rs:115-141
        while (number_of_bits > 0) {
            uint8_t bits_written;
            if (number_of_bits > remaining_bits) {
                unsigned shift = number_of_bits - remaining_bits;
                current_byte |= (uint8_t)((bits >> shift) & ((1u << remaining_bits) - 1));
                bits_written = remaining_bits;
            } else {
                unsigned shift = remaining_bits - number_of_bits;
                uint64_t mask = (uint64_t)(0xFFu >> (8 - remaining_bits));
                current_byte |= (uint8_t)((bits << shift) & mask);
                bits_written = number_of_bits;
            }
            number_of_bits -= bits_written;
            remaining_bits -= bits_written;
            if (remaining_bits == 0) {
                bytes.push_back(current_byte);
                current_byte = 0;
                remaining_bits = 8;
            }
        }
    }
    bool is_empty() const { return bytes.empty(); }                       // bits.rs:145-147 (bytes!)
    std::vector<uint8_t> finish() {                                       // bits.rs:157-162
        if (remaining_bits != 8) bytes.push_back(current_byte);
        return std::move(bytes);
    }
    std::vector<uint8_t> finish_with_one_bits() {                         // bits.rs:167-173
        if (remaining_bits != 8) {
            uint8_t remaining_bits_to_set = (uint8_t)((1u << remaining_bits) - 1);
            append_bits(remaining_bits_to_set, remaining_bits);
        }
        return finish();
    }
};

// ---------------------------------------------------------------------------------------------
// models/timestamps.rs (MacaqueTS)
// ---------------------------------------------------------------------------------------------

// timestamps.rs:77-95
bool are_uncompressed_timestamps_regular(const int64_t *ts, size_t n) {
    if (n < 2) return true;
    int64_t expected = ts[1] - ts[0];
    for (size_t i = 1; i < n; i++)
        if (ts[i] - ts[i - 1] != expected) return false;
    return true;
}

// timestamps.rs:99-108
std::vector<uint8_t> compress_regular_residual_timestamps(size_t length) {
    unsigned leading_zero_bits = length == 0 ? 64 : (unsigned)__builtin_clzll((unsigned long long)length);
    size_t number_of_bits_to_write = (64 - leading_zero_bits) + 1;
    size_t number_of_bytes_to_write = (size_t)std::ceil((double)number_of_bits_to_write / 8.0);
    std::vector<uint8_t> out(number_of_bytes_to_write);
    for (size_t i = 0; i < number_of_bytes_to_write; i++)
        out[number_of_bytes_to_write - 1 - i] = (uint8_t)(((uint64_t)length >> (8 * i)) & 0xFF);
    return out;
}

// timestamps.rs:113-155
std::vector<uint8_t> compress_irregular_residual_timestamps(const int64_t *ts, size_t n) {
    BitVecBuilder b;
    b.append_a_one_bit();
    int64_t last_timestamp = ts[0];
    int64_t last_delta = 0;
    for (size_t i = 1; i + 1 < n; i++) {
        // wrapping arithmetic: the reference is built in release mode for production.
        int64_t delta = (int64_t)((uint64_t)ts[i] - (uint64_t)last_timestamp);
        int64_t dod = (int64_t)((uint64_t)delta - (uint64_t)last_delta);
        if (dod == 0) {
            b.append_a_zero_bit();
        } else if (dod >= -63 && dod <= 64) {
            b.append_bits(0b10, 2);
            b.append_bits((uint64_t)dod, 7);
        } else if (dod >= -255 && dod <= 256) {
            b.append_bits(0b110, 3);
            b.append_bits((uint64_t)dod, 9);
        } else if (dod >= -2047 && dod <= 2048) {
            b.append_bits(0b1110, 4);
            b.append_bits((uint64_t)dod, 12);
        } else if (dod >= -2147483647LL && dod <= 2147483648LL) {
            b.append_bits(0b11110, 5);
            b.append_bits((uint64_t)dod, 32);
        } else {
            b.append_bits(0b11111, 5);
            b.append_bits((uint64_t)dod, 64);
        }
        last_delta = delta;
        last_timestamp = ts[i];
    }
    return b.finish_with_one_bits();
}

// timestamps.rs:56-73
std::vector<uint8_t> compress_residual_timestamps(const int64_t *ts, size_t n) {
    if (n <= 2) return {};
    if (are_uncompressed_timestamps_regular(ts, n)) return compress_regular_residual_timestamps(n);
    return compress_irregular_residual_timestamps(ts, n);
}

// timestamps.rs:199-202
inline bool are_compressed_timestamps_regular(const uint8_t *b, size_t n) {
    return n == 0 || (b[0] & 128) == 0;
}

inline uint64_t be_bytes_to_u64(const uint8_t *b, size_t n) {
    uint64_t v = 0;
    for (size_t i = 0; i < n && i < 8; i++) v = (v << 8) | b[i];
    return v;
}

// timestamps.rs:283-292
inline uint64_t read_decode_and_compute_delta(BitReader &bits, uint8_t bits_to_read, uint64_t last_delta) {
    uint64_t encoded = bits.read_bits(bits_to_read);
    uint64_t dod = encoded;
    if (encoded > ((uint64_t)1 << (bits_to_read - 1)))
        dod = encoded | (bits_to_read >= 64 ? 0 : (~(uint64_t)0 << bits_to_read));
    return last_delta + dod; // wrapping_add
}

// timestamps.rs:163-275.  Appends to `out`.
void decompress_all_timestamps(int64_t start_time, int64_t end_time, const uint8_t *b, size_t nb,
                               std::vector<int64_t> &out) {
    if (nb == 0 && start_time == end_time) {
        out.push_back(start_time);
    } else if (nb == 0) {
        out.push_back(start_time);
        out.push_back(end_time);
    } else if (are_compressed_timestamps_regular(b, nb)) {
        // timestamps.rs:207-223: (start..=end).step_by(interval)
        uint64_t length = be_bytes_to_u64(b, nb);
        if (length < 2) { out.push_back(start_time); return; } // malformed; the reference divides by zero
        uint64_t sampling_interval = (uint64_t)(end_time - start_time) / (length - 1);
        for (int64_t t = start_time; t <= end_time; t += (int64_t)sampling_interval) {
            out.push_back(t);
            if (sampling_interval == 0) break; // step_by(0) panics in the reference; never valid
        }
    } else {
        // timestamps.rs:228-275
        out.push_back(start_time);
        BitReader bits(b, nb);
        bits.read_bit();
        uint64_t last_delta = 0;
        int64_t timestamp = start_time;
        while (!bits.is_empty()) {
            int leading_one_bits = 0;
            while (leading_one_bits < 5 && !bits.is_empty() && bits.read_bit()) leading_one_bits++;
            if (leading_one_bits != 0 && bits.remaining_bits() < 7) break;
            uint64_t delta;
            switch (leading_one_bits) {
            case 0: delta = last_delta; break;
            case 1: delta = read_decode_and_compute_delta(bits, 7, last_delta); break;
            case 2: delta = read_decode_and_compute_delta(bits, 9, last_delta); break;
            case 3: delta = read_decode_and_compute_delta(bits, 12, last_delta); break;
            case 4: delta = read_decode_and_compute_delta(bits, 32, last_delta); break;
            default: delta = read_decode_and_compute_delta(bits, 64, last_delta); break;
            }
            timestamp = (int64_t)((uint64_t)timestamp + delta);
            out.push_back(timestamp);
            last_delta = delta;
        }
        out.push_back(end_time);
    }
}

// models/mod.rs:98-124
size_t segment_len(int64_t start_time, int64_t end_time, const uint8_t *b, size_t nb) {
    if (nb == 0 && start_time == end_time) return 1;
    if (nb == 0) return 2;
    if (are_compressed_timestamps_regular(b, nb)) return (size_t)be_bytes_to_u64(b, nb);
    std::vector<int64_t> tmp;
    decompress_all_timestamps(start_time, end_time, b, nb, tmp);
    return tmp.size();
}

// ---------------------------------------------------------------------------------------------
// models/pmc_mean.rs
// ---------------------------------------------------------------------------------------------

constexpr float COMPRESSED_METADATA_SIZE_IN_BYTES = 29.0f; // modelardb_types/src/schemas.rs:57-64
constexpr uint8_t VALUE_SIZE_IN_BYTES = 4;                 // models/mod.rs:47
constexpr uint8_t VALUE_SIZE_IN_BITS = 32;                 // models/mod.rs:50

struct PMCMean {                                           // pmc_mean.rs:31-53
    ErrorBound error_bound;
    float min_value = std::numeric_limits<float>::quiet_NaN();
    float max_value = std::numeric_limits<float>::quiet_NaN();
    double sum_of_values = 0.0;
    size_t length = 0;
    explicit PMCMean(ErrorBound eb) : error_bound(eb) {}
    bool fit_value(float value) {                          // pmc_mean.rs:58-75
        float next_min_value = rust_min(min_value, value);
        float next_max_value = rust_max(max_value, value);
        double next_sum_of_values = sum_of_values + (double)value;
        size_t next_length = length + 1;
        float average = (float)(next_sum_of_values / (double)next_length);
        if (is_value_within_error_bound(error_bound, next_min_value, average) &&
            is_value_within_error_bound(error_bound, next_max_value, average)) {
            min_value = next_min_value;
            max_value = next_max_value;
            sum_of_values = next_sum_of_values;
            length = next_length;
            return true;
        }
        return false;
    }
    float bytes_per_value() const { return COMPRESSED_METADATA_SIZE_IN_BYTES / (float)length; } // :83-87
    float model() const { return (float)(sum_of_values / (double)length); }                     // :91-93
};

inline float pmc_mean_sum(size_t model_length, float value) { return (float)model_length * value; } // :98-100

// ---------------------------------------------------------------------------------------------
// models/swing.rs
// ---------------------------------------------------------------------------------------------

// swing.rs:323-340
inline void compute_slope_and_intercept(int64_t start_time, double first_value, int64_t end_time,
                                        double last_value, double &slope, double &intercept) {
    if (equal_or_nan(first_value, last_value)) {
        slope = 0.0;
        intercept = first_value;
    } else {
        slope = (last_value - first_value) / (double)(end_time - start_time);
        intercept = first_value - slope * (double)start_time;
    }
}

struct Swing {                                             // swing.rs:34-80
    ErrorBound error_bound;
    int64_t start_time = 0, end_time = 0;
    double first_value = std::numeric_limits<double>::quiet_NaN();
    double upper_bound_slope = std::numeric_limits<double>::quiet_NaN();
    double upper_bound_intercept = std::numeric_limits<double>::quiet_NaN();
    double lower_bound_slope = std::numeric_limits<double>::quiet_NaN();
    double lower_bound_intercept = std::numeric_limits<double>::quiet_NaN();
    double mse_numerator = 0.0, mse_denominator = 0.0;
    size_t length = 0;
    explicit Swing(ErrorBound eb) : error_bound(eb) {}

    bool fit_data_point(int64_t timestamp, float value_f32) { // swing.rs:101-198
        double value = (double)value_f32;
        double maximum_deviation = maximum_allowed_deviation(error_bound, value);
        if (length == 0) {
            start_time = timestamp;
            end_time = timestamp;
            first_value = value;
            length += 1;
            return true;
        } else if (!std::isfinite(first_value) || !std::isfinite(value)) {
            if (equal_or_nan(first_value, value)) {
                end_time = timestamp;
                upper_bound_slope = value;
                upper_bound_intercept = value;
                lower_bound_slope = value;
                lower_bound_intercept = value;
                length += 1;
                return true;
            }
            return false;
        } else if (length == 1) {
            end_time = timestamp;
            compute_slope_and_intercept(start_time, first_value, timestamp, value + maximum_deviation,
                                        upper_bound_slope, upper_bound_intercept);
            compute_slope_and_intercept(start_time, first_value, timestamp, value - maximum_deviation,
                                        lower_bound_slope, lower_bound_intercept);
            length += 1;
            return true;
        } else {
            double upper = upper_bound_slope * (double)timestamp + upper_bound_intercept;
            double lower = lower_bound_slope * (double)timestamp + lower_bound_intercept;
            if (upper + maximum_deviation < value || lower - maximum_deviation > value) return false;
            end_time = timestamp;
            if (upper - maximum_deviation > value)
                compute_slope_and_intercept(start_time, first_value, timestamp, value + maximum_deviation,
                                            upper_bound_slope, upper_bound_intercept);
            if (lower + maximum_deviation < value)
                compute_slope_and_intercept(start_time, first_value, timestamp, value - maximum_deviation,
                                            lower_bound_slope, lower_bound_intercept);
            // swing.rs:212-228
            if (!equal_or_nan(first_value, value)) {
                double dt = (double)(timestamp - start_time);
                mse_numerator += (value - first_value) * dt;
                mse_denominator += dt * dt; // powi(2)
            } else {
                mse_numerator += 0.0;
                mse_denominator += 0.0;
            }
            length += 1;
            return true;
        }
    }
    float bytes_per_value() const { return (COMPRESSED_METADATA_SIZE_IN_BYTES + 1.0f) / (float)length; } // :236-239
    void model(float &first_out, float &last_out) const {   // swing.rs:246-259
        double projected_slope = mse_numerator / mse_denominator;
        double slope = rust_max(lower_bound_slope, rust_min(projected_slope, upper_bound_slope));
        double last_value = slope * (double)(end_time - start_time) + first_value;
        first_out = (float)first_value;
        last_out = (float)last_value;
    }
};

// ---------------------------------------------------------------------------------------------
// models/macaque_v.rs
// ---------------------------------------------------------------------------------------------

inline int32_t get_exponent(float value) { return (int32_t)((f32_bits(value) >> 23) & 0xff) - 127; } // :326-330

// macaque_v.rs:333-336 with release-mode shift semantics (the amount is masked to 5 bits).
inline uint32_t rewrite_bits_by_n(uint32_t bits_to_rewrite, int32_t positions_to_shift) {
    uint32_t mask = 0xFFFFFFFFu << ((uint32_t)positions_to_shift & 31u);
    return bits_to_rewrite & mask;
}

// `23 - factorized_epsilon.log2().abs().floor() as i32` (macaque_v.rs:185): f32::log2 is libm's log2f.
inline int32_t rewrite_position_libm(float factorized_epsilon) {
    return 23 - f32_as_i32(std::floor(std::fabs(log2f(factorized_epsilon))));
}
// What the GPU computes: log2 in f64, rounded once to f32.  tests/test_oracle_log2.py checks it
// equals the libm form for every f32 input.
inline int32_t rewrite_position_f64(float factorized_epsilon) {
    float l = (float)std::log2((double)factorized_epsilon);
    return 23 - f32_as_i32(std::floor(std::fabs(l)));
}

// 2f32.powi(e): compiler-rt __powisf2 by repeated squaring; exact for base 2 (incl. the subnormal
// 2^-127), computed the same way here.
inline float powi2(int32_t b) {
    const bool recip = b < 0;
    float a = 2.0f, r = 1.0f;
    while (true) {
        if (b & 1) r *= a;
        b /= 2;
        if (b == 0) break;
        a *= a;
    }
    return recip ? 1.0f / r : r;
}

struct MacaqueV {                                          // macaque_v.rs:39-72
    ErrorBound error_bound;
    float min_value = std::numeric_limits<float>::quiet_NaN();
    float max_value = std::numeric_limits<float>::quiet_NaN();
    float last_value = 0.0f;
    uint8_t last_leading_zero_bits = 255;
    uint8_t last_trailing_zero_bits = 0;
    BitVecBuilder compressed_values;
    size_t length = 0;
    explicit MacaqueV(ErrorBound eb) : error_bound(eb) {}

    void update_min_max_and_last_value(float value) {      // :199-204
        min_value = rust_min(min_value, value);
        max_value = rust_max(max_value, value);
        last_value = value;
        length += 1;
    }
    void compress_values(const float *values, size_t n) {  // :76-88
        for (size_t i = 0; i < n; i++) {
            if (compressed_values.is_empty()) {
                compressed_values.append_bits((uint64_t)f32_bits(values[i]), VALUE_SIZE_IN_BITS);
                update_min_max_and_last_value(values[i]);
            } else {
                compress_value_xor_last_value(values[i]);
            }
        }
    }
    void compress_values_without_first(const float *values, size_t n, float model_last_value) { // :92-97
        last_value = model_last_value;
        for (size_t i = 0; i < n; i++) compress_value_xor_last_value(values[i]);
    }
    float rewrite_least_mantissa_bits(float value) const { // :168-196
        if (std::fabs(value) == 0.0f || std::isnan(value) || std::isinf(value)) return value;
        uint32_t value_as_u32 = f32_bits(value);
        float abs_error_bound = (float)maximum_allowed_deviation(error_bound, (double)value);
        int32_t exponent = get_exponent(value);
        float factorized_epsilon = abs_error_bound / powi2(exponent);
        int32_t rewrite_position = rewrite_position_libm(factorized_epsilon);
        float rewritten_value = f32_from_bits(rewrite_bits_by_n(value_as_u32, rewrite_position));
        if (!is_value_within_error_bound(error_bound, value, rewritten_value)) {
            rewrite_position -= 1;
            rewritten_value = f32_from_bits(rewrite_bits_by_n(value_as_u32, rewrite_position));
        }
        return rewritten_value;
    }
    void compress_value_xor_last_value(float value) {      // :100-164
        if (error_bound.kind != MDBO_LOSSLESS) {
            if (is_value_within_error_bound(error_bound, value, last_value)) value = last_value;
            else value = rewrite_least_mantissa_bits(value);
        }
        uint32_t x = f32_bits(value) ^ f32_bits(last_value);
        if (x == 0) {
            compressed_values.append_a_one_bit();
            compressed_values.append_a_zero_bit();
        } else {
            uint8_t lz = (uint8_t)__builtin_clz(x);
            uint8_t tz = (uint8_t)__builtin_ctz(x);
            if (lz >= last_leading_zero_bits && tz >= last_trailing_zero_bits) {
                compressed_values.append_a_zero_bit();
                uint8_t meaningful_bits = VALUE_SIZE_IN_BITS - last_leading_zero_bits - last_trailing_zero_bits;
                compressed_values.append_bits((uint64_t)(x >> last_trailing_zero_bits), meaningful_bits);
            } else {
                compressed_values.append_a_one_bit();
                compressed_values.append_a_one_bit();
                compressed_values.append_bits((uint64_t)lz, 5);
                uint8_t meaningful_bits = VALUE_SIZE_IN_BITS - lz - tz;
                compressed_values.append_bits((uint64_t)meaningful_bits, 6);
                compressed_values.append_bits((uint64_t)(x >> tz), meaningful_bits);
                last_leading_zero_bits = lz;
                last_trailing_zero_bits = tz;
            }
        }
        update_min_max_and_last_value(value);
    }
};

// macaque_v.rs:272-323 (grid) and :220-265 (sum) share this decoder; Sink gets each decoded value.
template <typename Sink>
void macaque_v_decode(const uint8_t *values, size_t n_bytes, size_t n_values, bool has_seed, float seed,
                      Sink &&sink) {
    if (n_values == 0) return; // the reference underflows `length - 1` here; never produced by the encoder
    BitReader bits(values, n_bytes);
    uint8_t leading_zeros = 255;
    uint8_t trailing_zeros = 0;
    uint32_t last_value;
    if (has_seed) {
        last_value = f32_bits(seed);
    } else {
        last_value = (uint32_t)bits.read_bits(VALUE_SIZE_IN_BITS);
        sink(f32_from_bits(last_value));
    }
    size_t remaining = n_values - (has_seed ? 0 : 1);
    for (size_t i = 0; i < remaining; i++) {
        if (bits.read_bit()) {
            if (bits.read_bit()) {
                leading_zeros = (uint8_t)bits.read_bits(5);
                uint8_t meaningful_bits = (uint8_t)bits.read_bits(6);
                trailing_zeros = (uint8_t)(VALUE_SIZE_IN_BITS - meaningful_bits - leading_zeros);
                meaningful_bits = (uint8_t)(VALUE_SIZE_IN_BITS - leading_zeros - trailing_zeros);
                uint32_t value = (uint32_t)bits.read_bits(meaningful_bits);
                value <<= trailing_zeros;
                value ^= last_value;
                last_value = value;
            }
        } else {
            uint8_t meaningful_bits = (uint8_t)(VALUE_SIZE_IN_BITS - leading_zeros - trailing_zeros);
            uint32_t value = (uint32_t)bits.read_bits(meaningful_bits);
            value <<= trailing_zeros;
            value ^= last_value;
            last_value = value;
        }
        sink(f32_from_bits(last_value));
    }
}

float macaque_v_sum(size_t length, const uint8_t *values, size_t n_bytes, bool has_seed, float seed) {
    float sum = 0.0f; // macaque_v.rs:228-235: (0.0 | first value) then += in order, all f32
    bool first = true;
    macaque_v_decode(values, n_bytes, length, has_seed, seed, [&](float v) {
        if (first && !has_seed) sum = v; else sum += v;
        first = false;
    });
    return sum;
}

// ---------------------------------------------------------------------------------------------
// types.rs: ModelBuilder, CompressedSegmentBuilder
// ---------------------------------------------------------------------------------------------

struct SegmentBatch { // types.rs:411-517, schema modelardb_types/src/schemas.rs:40-52 (error is always NaN)
    std::vector<int8_t> model_type_id;
    std::vector<int64_t> start_time, end_time;
    std::vector<float> min_value, max_value;
    std::vector<uint64_t> timestamps_off{0}, values_off{0}, residuals_off{0};
    std::vector<uint8_t> timestamps_data, values_data, residuals_data;
    void append(int8_t id, int64_t st, int64_t et, const std::vector<uint8_t> &ts, float mn, float mx,
                const uint8_t *vals, size_t n_vals, const std::vector<uint8_t> &res) { // types.rs:468-489
        model_type_id.push_back(id);
        start_time.push_back(st);
        end_time.push_back(et);
        timestamps_data.insert(timestamps_data.end(), ts.begin(), ts.end());
        timestamps_off.push_back(timestamps_data.size());
        min_value.push_back(mn);
        max_value.push_back(mx);
        values_data.insert(values_data.end(), vals, vals + n_vals);
        values_off.push_back(values_data.size());
        residuals_data.insert(residuals_data.end(), res.begin(), res.end());
        residuals_off.push_back(residuals_data.size());
    }
    void append_batch(const SegmentBatch &o) {
        size_t n = o.model_type_id.size();
        model_type_id.insert(model_type_id.end(), o.model_type_id.begin(), o.model_type_id.end());
        start_time.insert(start_time.end(), o.start_time.begin(), o.start_time.end());
        end_time.insert(end_time.end(), o.end_time.begin(), o.end_time.end());
        min_value.insert(min_value.end(), o.min_value.begin(), o.min_value.end());
        max_value.insert(max_value.end(), o.max_value.begin(), o.max_value.end());
        uint64_t tb = timestamps_data.size(), vb = values_data.size(), rb = residuals_data.size();
        for (size_t i = 1; i <= n; i++) {
            timestamps_off.push_back(tb + o.timestamps_off[i]);
            values_off.push_back(vb + o.values_off[i]);
            residuals_off.push_back(rb + o.residuals_off[i]);
        }
        timestamps_data.insert(timestamps_data.end(), o.timestamps_data.begin(), o.timestamps_data.end());
        values_data.insert(values_data.end(), o.values_data.begin(), o.values_data.end());
        residuals_data.insert(residuals_data.end(), o.residuals_data.begin(), o.residuals_data.end());
    }
};

struct CompressedSegmentBuilder { // types.rs:148-166
    int8_t model_type_id;
    size_t start_index, end_index;
    float min_value, max_value;
    std::vector<uint8_t> values;
    float model_last_value;
    float bytes_per_value;
    size_t pmc_len = 0, swing_len = 0;
};

// types.rs:283-303
std::vector<uint8_t> encode_values_for_pmc_mean(float min_value, float max_value, float residuals_min_value,
                                                float residuals_max_value) {
    std::vector<uint8_t> values;
    if (min_value > residuals_min_value) {
        if (max_value >= residuals_max_value) {
            values.push_back(1);
        } else {
            uint32_t b = f32_bits(min_value);
            for (int i = 0; i < 4; i++) values.push_back((uint8_t)(b >> (8 * i)));
        }
    }
    return values;
}

inline void push_le(std::vector<uint8_t> &v, float f) {
    uint32_t b = f32_bits(f);
    for (int i = 0; i < 4; i++) v.push_back((uint8_t)(b >> (8 * i)));
}

// types.rs:325-370
std::vector<uint8_t> encode_values_for_swing(float min_value, float max_value, bool min_value_is_first,
                                             float residuals_min_value, float residuals_max_value) {
    std::vector<uint8_t> values;
    if (residuals_min_value < min_value && max_value < residuals_max_value) {
        if (min_value_is_first) { push_le(values, min_value); push_le(values, max_value); }
        else { push_le(values, max_value); push_le(values, min_value); }
    } else if (residuals_min_value < min_value) {
        values.push_back(min_value_is_first ? 0 : 1);
        push_le(values, min_value);
    } else if (max_value < residuals_max_value) {
        values.push_back(min_value_is_first ? 2 : 3);
        push_le(values, max_value);
    } else if (!min_value_is_first) {
        values.push_back(0);
    }
    return values;
}

inline float le_f32(const uint8_t *p) {
    return f32_from_bits((uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24));
}

// types.rs:307-321.  The reference panics unless n is 0, 1 or 4; callers validate first.
inline float decode_values_for_pmc_mean(float min_value, float max_value, const uint8_t *values, size_t n) {
    if (n == 0) return min_value;
    if (n == 1) return max_value;
    return le_f32(values);
}

// types.rs:374-407.  Returns false where the reference panics.
inline bool decode_values_for_swing(float min_value, float max_value, const uint8_t *values, size_t n,
                                    float &first, float &last) {
    switch (n) {
    case 0: first = min_value; last = max_value; return true;
    case 1: first = max_value; last = min_value; return true;
    case 5: {
        float value = le_f32(values + 1);
        switch (values[0]) {
        case 0: first = value; last = max_value; return true;
        case 1: first = max_value; last = value; return true;
        case 2: first = min_value; last = value; return true;
        case 3: first = value; last = min_value; return true;
        default: return false;
        }
    }
    case 8: first = le_f32(values); last = le_f32(values + 4); return true;
    default: return false;
    }
}

// types.rs:104-144
CompressedSegmentBuilder select_pmc_mean(size_t start_index, const PMCMean &pmc) {
    float value = pmc.model();
    return {MDBO_PMC_MEAN, start_index, start_index + pmc.length - 1, value, value, {}, value,
            pmc.bytes_per_value()};
}
CompressedSegmentBuilder select_swing(size_t start_index, const Swing &swing) {
    float first_value, last_value;
    swing.model(first_value, last_value);
    float min_value = rust_min(first_value, last_value);
    float max_value = rust_max(first_value, last_value);
    std::vector<uint8_t> values;
    if (!(first_value < last_value)) values.push_back(0);
    return {MDBO_SWING, start_index, start_index + swing.length - 1, min_value, max_value, values,
            last_value, swing.bytes_per_value()};
}

// compression.rs:280-301 + types.rs:61-101
CompressedSegmentBuilder fit_next_model(size_t current_start_index, ErrorBound eb, const int64_t *ts,
                                        const float *values, size_t end_index) {
    PMCMean pmc(eb);
    Swing swing(eb);
    bool pmc_could_fit_all = true, swing_could_fit_all = true;
    size_t current_index = current_start_index;
    bool can_fit_more = true;
    while (can_fit_more && current_index < end_index) {
        int64_t timestamp = ts[current_index];
        float value = values[current_index];
        pmc_could_fit_all = pmc_could_fit_all && pmc.fit_value(value);          // types.rs:75
        swing_could_fit_all = swing_could_fit_all && swing.fit_data_point(timestamp, value); // :77-78
        can_fit_more = pmc_could_fit_all || swing_could_fit_all;
        current_index += 1;
    }
    // types.rs:84-101: min_by returns the FIRST minimum, so PMC-Mean wins ties.
    float pmc_bpv = pmc.bytes_per_value(), swing_bpv = swing.bytes_per_value();
    CompressedSegmentBuilder b = (swing_bpv < pmc_bpv) ? select_swing(current_start_index, swing)
                                                      : select_pmc_mean(current_start_index, pmc);
    b.pmc_len = pmc.length;
    b.swing_len = swing.length;
    return b;
}

// types.rs:197-267
void segment_finish(CompressedSegmentBuilder self, ErrorBound eb, size_t residuals_end_index,
                    const int64_t *ts, const float *values, SegmentBatch &out) {
    int64_t start_time = ts[self.start_index];
    int64_t end_time = ts[residuals_end_index];
    std::vector<uint8_t> timestamps =
        compress_residual_timestamps(ts + self.start_index, residuals_end_index - self.start_index + 1);
    std::vector<uint8_t> residuals;
    if (self.end_index < residuals_end_index) {
        size_t residuals_start_index = self.end_index + 1;
        MacaqueV macaque_v(eb); // types.rs:270-278
        macaque_v.compress_values_without_first(values + residuals_start_index,
                                                residuals_end_index - residuals_start_index + 1,
                                                self.model_last_value);
        float residuals_min_value = macaque_v.min_value, residuals_max_value = macaque_v.max_value;
        residuals = macaque_v.compressed_values.finish();
        if (self.model_type_id == MDBO_PMC_MEAN)
            self.values = encode_values_for_pmc_mean(self.min_value, self.max_value, residuals_min_value,
                                                     residuals_max_value);
        else
            self.values = encode_values_for_swing(self.min_value, self.max_value, self.values.empty(),
                                                  residuals_min_value, residuals_max_value);
        self.min_value = rust_min(self.min_value, residuals_min_value);
        self.max_value = rust_max(self.max_value, residuals_max_value);
        residuals.push_back((uint8_t)((residuals_end_index - residuals_start_index) + 1));
    }
    out.append(self.model_type_id, start_time, end_time, timestamps, self.min_value, self.max_value,
               self.values.data(), self.values.size(), residuals);
}

// compression.rs:367-400
void compress_and_store_residuals_in_a_separate_segment(ErrorBound eb, size_t start_index, size_t end_index,
                                                        const int64_t *ts, const float *values,
                                                        SegmentBatch &out) {
    std::vector<uint8_t> timestamps = compress_residual_timestamps(ts + start_index, end_index - start_index + 1);
    MacaqueV macaque_v(eb);
    macaque_v.compress_values(values + start_index, end_index - start_index + 1);
    float mn = macaque_v.min_value, mx = macaque_v.max_value;
    std::vector<uint8_t> bytes = macaque_v.compressed_values.finish();
    out.append(MDBO_MACAQUE_V, ts[start_index], ts[end_index], timestamps, mn, mx, bytes.data(), bytes.size(), {});
}

constexpr size_t RESIDUAL_VALUES_MAX_LENGTH = 255; // compression.rs:38

// compression.rs:310-362
void store_compressed_segments_with_model_and_or_residuals(ErrorBound eb, const CompressedSegmentBuilder *maybe_model,
                                                           size_t residuals_end_index, const int64_t *ts,
                                                           const float *values, SegmentBatch &out) {
    if (maybe_model) {
        if ((residuals_end_index - maybe_model->end_index) <= RESIDUAL_VALUES_MAX_LENGTH) {
            segment_finish(*maybe_model, eb, residuals_end_index, ts, values, out);
        } else {
            size_t model_end_index = maybe_model->end_index;
            segment_finish(*maybe_model, eb, model_end_index, ts, values, out);
            compress_and_store_residuals_in_a_separate_segment(eb, model_end_index + 1, residuals_end_index, ts,
                                                               values, out);
        }
    } else {
        compress_and_store_residuals_in_a_separate_segment(eb, 0, residuals_end_index, ts, values, out);
    }
}

// compression.rs:191-275
void try_compress_univariate_time_series(const int64_t *ts, const float *values, size_t end_index, ErrorBound eb,
                                         SegmentBatch &out) {
    if (end_index == 0) return;
    size_t current_start_index = 0;
    bool have_previous = false;
    CompressedSegmentBuilder previous_model{};
    while (current_start_index < end_index) {
        CompressedSegmentBuilder model = fit_next_model(current_start_index, eb, ts, values, end_index);
        if (model.bytes_per_value <= (float)VALUE_SIZE_IN_BYTES) {
            if (current_start_index > 0)
                store_compressed_segments_with_model_and_or_residuals(eb, have_previous ? &previous_model : nullptr,
                                                                      current_start_index - 1, ts, values, out);
            current_start_index = model.end_index + 1;
            previous_model = std::move(model);
            have_previous = true;
        } else {
            current_start_index += 1;
        }
    }
    store_compressed_segments_with_model_and_or_residuals(eb, have_previous ? &previous_model : nullptr,
                                                          end_index - 1, ts, values, out);
}

// ---------------------------------------------------------------------------------------------
// models/mod.rs: grid / sum
// ---------------------------------------------------------------------------------------------

inline size_t residuals_length(const uint8_t *residuals, size_t n) { return n == 0 ? 0 : residuals[n - 1]; } // :277-284

// swing.rs:264-300
float swing_sum(int64_t start_time, int64_t end_time, const uint8_t *tsb, size_t n_tsb, float first_value,
                float last_value, size_t residuals_len) {
    double slope, intercept;
    compute_slope_and_intercept(start_time, (double)first_value, end_time, (double)last_value, slope, intercept);
    if (are_compressed_timestamps_regular(tsb, n_tsb)) {
        double first = slope * (double)start_time + intercept;
        double last = slope * (double)end_time + intercept;
        double average = (first + last) / 2.0;
        size_t length = segment_len(start_time, end_time, tsb, n_tsb) - residuals_len;
        return (float)(average * (double)length);
    }
    std::vector<int64_t> timestamps;
    decompress_all_timestamps(start_time, end_time, tsb, n_tsb, timestamps);
    size_t model_timestamps_end_index = timestamps.size() - residuals_len;
    double sum = 0.0;
    for (size_t i = 0; i < model_timestamps_end_index; i++) sum += slope * (double)timestamps[i] + intercept;
    return (float)sum;
}

struct Row {
    int8_t model_type_id;
    int64_t start_time, end_time;
    const uint8_t *timestamps; size_t n_timestamps;
    float min_value, max_value;
    const uint8_t *values; size_t n_values;
    const uint8_t *residuals; size_t n_residuals;
};

inline Row row_of(const mdbo_segments_view *v, uint64_t i) {
    return {v->model_type_id[i], v->start_time[i], v->end_time[i],
            v->timestamps_data + v->timestamps_off[i], (size_t)(v->timestamps_off[i + 1] - v->timestamps_off[i]),
            v->min_value[i], v->max_value[i],
            v->values_data + v->values_off[i], (size_t)(v->values_off[i + 1] - v->values_off[i]),
            v->residuals_data + v->residuals_off[i], (size_t)(v->residuals_off[i + 1] - v->residuals_off[i])};
}

// models/mod.rs:129-184
float segment_sum(const Row &r) {
    size_t res_len = residuals_length(r.residuals, r.n_residuals);
    size_t model_length = segment_len(r.start_time, r.end_time, r.timestamps, r.n_timestamps) - res_len;
    float model_last_value, model_sum;
    switch (r.model_type_id) {
    case MDBO_PMC_MEAN: {
        float value = decode_values_for_pmc_mean(r.min_value, r.max_value, r.values, r.n_values);
        model_last_value = value;
        model_sum = pmc_mean_sum(model_length, value);
        break;
    }
    case MDBO_SWING: {
        float first_value = 0, last_value = 0;
        decode_values_for_swing(r.min_value, r.max_value, r.values, r.n_values, first_value, last_value);
        model_last_value = last_value;
        model_sum = swing_sum(r.start_time, r.end_time, r.timestamps, r.n_timestamps, first_value, last_value, res_len);
        break;
    }
    default:
        model_last_value = std::numeric_limits<float>::quiet_NaN();
        model_sum = macaque_v_sum(model_length, r.values, r.n_values, false, 0.0f);
        break;
    }
    if (r.n_residuals == 0) return model_sum;
    float residuals_sum = macaque_v_sum(res_len, r.residuals, r.n_residuals - 1, true, model_last_value);
    return model_sum + residuals_sum;
}

// models/mod.rs:190-251.  Appends to ts_out/val_out.
void segment_grid(const Row &r, std::vector<int64_t> &ts_out, std::vector<float> &val_out) {
    size_t res_len = residuals_length(r.residuals, r.n_residuals);
    size_t ts_begin = ts_out.size();
    decompress_all_timestamps(r.start_time, r.end_time, r.timestamps, r.n_timestamps, ts_out);
    size_t model_end = ts_out.size() - res_len; // absolute index one past the model's last timestamp
    size_t n_model = model_end - ts_begin;
    switch (r.model_type_id) {
    case MDBO_PMC_MEAN: { // pmc_mean.rs:104-108
        float value = decode_values_for_pmc_mean(r.min_value, r.max_value, r.values, r.n_values);
        for (size_t i = 0; i < n_model; i++) val_out.push_back(value);
        break;
    }
    case MDBO_SWING: {    // swing.rs:304-319 with model_end_time = last MODEL timestamp (mod.rs:223-234)
        float first_value = 0, last_value = 0;
        decode_values_for_swing(r.min_value, r.max_value, r.values, r.n_values, first_value, last_value);
        int64_t model_end_time = ts_out[model_end - 1];
        double slope, intercept;
        compute_slope_and_intercept(r.start_time, (double)first_value, model_end_time, (double)last_value, slope, intercept);
        for (size_t i = ts_begin; i < model_end; i++) val_out.push_back((float)(slope * (double)ts_out[i] + intercept));
        break;
    }
    default:
        macaque_v_decode(r.values, r.n_values, n_model, false, 0.0f, [&](float v) { val_out.push_back(v); });
        break;
    }
    if (r.n_residuals != 0) {
        float model_last_value = val_out.back(); // mod.rs:241-249: the last GRIDDED value (quirk Q1)
        macaque_v_decode(r.residuals, r.n_residuals - 1, res_len, true, model_last_value,
                         [&](float v) { val_out.push_back(v); });
    }
}

template <typename F> void parallel_for(uint64_t n, int n_threads, F &&f) {
    if (n_threads <= 1 || n < 2) { f(0, n, 0); return; }
    uint64_t nt = std::min<uint64_t>((uint64_t)n_threads, n);
    std::vector<std::thread> threads;
    for (uint64_t t = 0; t < nt; t++) {
        uint64_t lo = n * t / nt, hi = n * (t + 1) / nt;
        threads.emplace_back([=, &f] { f(lo, hi, (int)t); });
    }
    for (auto &th : threads) th.join();
}

} // namespace

struct mdbo_segments { SegmentBatch b; };

extern "C" {

int mdbo_is_value_within_error_bound(int kind, float eb, float real_value, float approximate_value) {
    return is_value_within_error_bound({kind, eb}, real_value, approximate_value) ? 1 : 0;
}
double mdbo_maximum_allowed_deviation(int kind, float eb, double value) {
    return maximum_allowed_deviation({kind, eb}, value);
}

size_t mdbo_compress_residual_timestamps(const int64_t *ts, size_t n, uint8_t *out, size_t cap) {
    std::vector<uint8_t> b = compress_residual_timestamps(ts, n);
    if (b.size() > cap) return (size_t)-1;
    if (!b.empty()) std::memcpy(out, b.data(), b.size());
    return b.size();
}
size_t mdbo_decompress_all_timestamps(int64_t start_time, int64_t end_time, const uint8_t *bytes, size_t n_bytes,
                                      int64_t *out, size_t cap) {
    std::vector<int64_t> v;
    decompress_all_timestamps(start_time, end_time, bytes, n_bytes, v);
    if (v.size() > cap) return (size_t)-1;
    if (!v.empty()) std::memcpy(out, v.data(), v.size() * 8);
    return v.size();
}
size_t mdbo_len(int64_t start_time, int64_t end_time, const uint8_t *bytes, size_t n_bytes) {
    return segment_len(start_time, end_time, bytes, n_bytes);
}

size_t mdbo_macaque_v_compress(int kind, float eb, const float *values, size_t n, int has_seed, float seed,
                               uint8_t *out, size_t cap, float *min_out, float *max_out, uint8_t *state_out,
                               float *last_value_out) {
    MacaqueV m({kind, eb});
    if (has_seed) m.compress_values_without_first(values, n, seed);
    else m.compress_values(values, n);
    if (min_out) *min_out = m.min_value;
    if (max_out) *max_out = m.max_value;
    if (state_out) { state_out[0] = m.last_leading_zero_bits; state_out[1] = m.last_trailing_zero_bits; }
    if (last_value_out) *last_value_out = m.last_value;
    std::vector<uint8_t> b = m.compressed_values.finish();
    if (b.size() > cap) return (size_t)-1;
    if (!b.empty()) std::memcpy(out, b.data(), b.size());
    return b.size();
}
void mdbo_macaque_v_grid(const uint8_t *bytes, size_t n_bytes, size_t n_values, int has_seed, float seed, float *out) {
    size_t i = 0;
    macaque_v_decode(bytes, n_bytes, n_values, has_seed != 0, seed, [&](float v) { out[i++] = v; });
}
float mdbo_macaque_v_sum(const uint8_t *bytes, size_t n_bytes, size_t n_values, int has_seed, float seed) {
    return macaque_v_sum(n_values, bytes, n_bytes, has_seed != 0, seed);
}
float mdbo_rewrite_least_mantissa_bits(int kind, float eb, float value) {
    MacaqueV m({kind, eb});
    return m.rewrite_least_mantissa_bits(value);
}
int32_t mdbo_rewrite_position_libm(float e) { return rewrite_position_libm(e); }
int32_t mdbo_rewrite_position_f64(float e) { return rewrite_position_f64(e); }

// rewrite_position as a step function of the bit pattern: every bit pattern b in [first_bits, last_bits] at which the
// position differs from the position at b - 1 (first_bits itself is always listed), in ascending order.  Scans every
// pattern of the range with `which` = 0 (libm log2f, the reference) or 1 (f64 log2 rounded once, the GPU's form) on
// n_threads threads; two step functions with the same list are the same function on the whole range.
size_t mdbo_rewrite_position_steps(int which, uint32_t first_bits, uint32_t last_bits, int n_threads, uint32_t *bits_out,
                                   int32_t *pos_out, size_t cap) {
    if (last_bits < first_bits) return 0;
    if (n_threads < 1) n_threads = 1;
    auto pos_of = [which](uint32_t b) { return which ? rewrite_position_f64(f32_from_bits(b)) : rewrite_position_libm(f32_from_bits(b)); };
    const uint64_t total = (uint64_t)last_bits - first_bits + 1;
    std::vector<std::vector<std::pair<uint32_t, int32_t>>> found((size_t)n_threads);
    std::vector<std::thread> threads;
    for (int t = 0; t < n_threads; t++) {
        threads.emplace_back([&, t] {
            const uint64_t lo = first_bits + total * (uint64_t)t / (uint64_t)n_threads, hi = first_bits + total * (uint64_t)(t + 1) / (uint64_t)n_threads;
            if (lo >= hi) return;
            int32_t prev = lo == first_bits ? pos_of((uint32_t)lo) + 1 : pos_of((uint32_t)(lo - 1)); // (first_bits is always a step)
            for (uint64_t b = lo; b < hi; b++) {
                const int32_t p = pos_of((uint32_t)b);
                if (p != prev) found[(size_t)t].emplace_back((uint32_t)b, p);
                prev = p;
            }
        });
    }
    for (auto &th : threads) th.join();
    size_t n = 0;
    for (auto &f : found)
        for (auto &e : f) {
            if (n < cap) { bits_out[n] = e.first; pos_out[n] = e.second; }
            n++;
        }
    return n;
}

static void model_to_c(const CompressedSegmentBuilder &b, mdbo_model *out) {
    std::memset(out, 0, sizeof(*out));
    out->model_type_id = b.model_type_id;
    out->start_index = b.start_index;
    out->end_index = b.end_index;
    out->min_value = b.min_value;
    out->max_value = b.max_value;
    out->values_len = (uint32_t)b.values.size();
    for (size_t i = 0; i < b.values.size() && i < 8; i++) out->values[i] = b.values[i];
    out->model_last_value = b.model_last_value;
    out->bytes_per_value = b.bytes_per_value;
    out->pmc_len = b.pmc_len;
    out->swing_len = b.swing_len;
}
static CompressedSegmentBuilder model_from_c(const mdbo_model *m) {
    CompressedSegmentBuilder b{m->model_type_id, (size_t)m->start_index, (size_t)m->end_index, m->min_value,
                               m->max_value, std::vector<uint8_t>(m->values, m->values + m->values_len),
                               m->model_last_value, m->bytes_per_value};
    return b;
}

void mdbo_fit_next_model(uint64_t start_index, int kind, float eb, const int64_t *ts, const float *values, uint64_t n,
                         mdbo_model *out) {
    model_to_c(fit_next_model((size_t)start_index, {kind, eb}, ts, values, (size_t)n), out);
}

uint64_t mdbo_pmc_fit_prefix(int kind, float eb, const float *values, uint64_t n, float *mean_out) {
    PMCMean pmc({kind, eb});
    for (uint64_t i = 0; i < n; i++)
        if (!pmc.fit_value(values[i])) break;
    if (mean_out) *mean_out = pmc.length ? pmc.model() : std::numeric_limits<float>::quiet_NaN();
    return pmc.length;
}
uint64_t mdbo_swing_fit_prefix(int kind, float eb, const int64_t *ts, const float *values, uint64_t n, float *first_out,
                               float *last_out) {
    Swing swing({kind, eb});
    for (uint64_t i = 0; i < n; i++)
        if (!swing.fit_data_point(ts[i], values[i])) break;
    float f, l;
    swing.model(f, l);
    if (first_out) *first_out = f;
    if (last_out) *last_out = l;
    return swing.length;
}

void mdbo_swing_bounds(int kind, float eb, const int64_t *ts, const float *values, uint64_t n, double *out4) {
    Swing swing({kind, eb});
    for (uint64_t i = 0; i < n; i++)
        if (!swing.fit_data_point(ts[i], values[i])) break;
    out4[0] = swing.lower_bound_slope; out4[1] = swing.lower_bound_intercept;
    out4[2] = swing.upper_bound_slope; out4[3] = swing.upper_bound_intercept;
}

mdbo_segments *mdbo_compress(const int64_t *ts, const float *values, const uint64_t *unit_off, uint64_t n_units,
                             const uint8_t *eb_kind, const float *eb_value, int n_threads, uint64_t *unit_seg_off_out) {
    std::vector<SegmentBatch> per_unit(n_units);
    parallel_for(n_units, n_threads, [&](uint64_t lo, uint64_t hi, int) {
        for (uint64_t u = lo; u < hi; u++)
            try_compress_univariate_time_series(ts + unit_off[u], values + unit_off[u],
                                                (size_t)(unit_off[u + 1] - unit_off[u]),
                                                {(int)eb_kind[u], eb_value[u]}, per_unit[u]);
    });
    mdbo_segments *s = new mdbo_segments();
    if (unit_seg_off_out) unit_seg_off_out[0] = 0;
    for (uint64_t u = 0; u < n_units; u++) {
        s->b.append_batch(per_unit[u]);
        if (unit_seg_off_out) unit_seg_off_out[u + 1] = s->b.model_type_id.size();
        per_unit[u] = SegmentBatch();
    }
    return s;
}
void mdbo_segments_view_get(const mdbo_segments *s, mdbo_segments_view *out) {
    const SegmentBatch &b = s->b;
    out->n_segments = b.model_type_id.size();
    out->model_type_id = b.model_type_id.data();
    out->start_time = b.start_time.data();
    out->end_time = b.end_time.data();
    out->min_value = b.min_value.data();
    out->max_value = b.max_value.data();
    out->timestamps_off = b.timestamps_off.data();
    out->timestamps_data = b.timestamps_data.data();
    out->values_off = b.values_off.data();
    out->values_data = b.values_data.data();
    out->residuals_off = b.residuals_off.data();
    out->residuals_data = b.residuals_data.data();
}
void mdbo_segments_free(mdbo_segments *s) { delete s; }

mdbo_segments *mdbo_model_finish(const mdbo_model *model, int kind, float eb, uint64_t residuals_end_index,
                                 const int64_t *ts, const float *values) {
    mdbo_segments *s = new mdbo_segments();
    segment_finish(model_from_c(model), {kind, eb}, (size_t)residuals_end_index, ts, values, s->b);
    return s;
}
mdbo_segments *mdbo_macaque_v_segment(int kind, float eb, uint64_t start_index, uint64_t end_index, const int64_t *ts,
                                      const float *values) {
    mdbo_segments *s = new mdbo_segments();
    compress_and_store_residuals_in_a_separate_segment({kind, eb}, (size_t)start_index, (size_t)end_index, ts, values, s->b);
    return s;
}

float mdbo_decode_values_for_pmc_mean(float min_value, float max_value, const uint8_t *values, size_t n) {
    return decode_values_for_pmc_mean(min_value, max_value, values, n);
}
int mdbo_decode_values_for_swing(float min_value, float max_value, const uint8_t *values, size_t n, float *first_out,
                                 float *last_out) {
    return decode_values_for_swing(min_value, max_value, values, n, *first_out, *last_out) ? 0 : 1;
}

uint64_t mdbo_grid_count(const mdbo_segments_view *v, uint64_t *point_off_out, int n_threads) {
    std::vector<uint64_t> lens(v->n_segments);
    parallel_for(v->n_segments, n_threads, [&](uint64_t lo, uint64_t hi, int) {
        for (uint64_t i = lo; i < hi; i++) {
            Row r = row_of(v, i);
            lens[i] = segment_len(r.start_time, r.end_time, r.timestamps, r.n_timestamps);
        }
    });
    uint64_t total = 0;
    for (uint64_t i = 0; i < v->n_segments; i++) {
        if (point_off_out) point_off_out[i] = total;
        total += lens[i];
    }
    if (point_off_out) point_off_out[v->n_segments] = total;
    return total;
}

uint64_t mdbo_grid(const mdbo_segments_view *v, int64_t *ts_out, float *val_out, uint64_t capacity, int n_threads) {
    // GridStream (crates/modelardb_storage/src/query/grid_exec.rs:323-337): rows in order, appended.
    std::vector<uint64_t> point_off(v->n_segments + 1);
    uint64_t total = mdbo_grid_count(v, point_off.data(), n_threads);
    if (total > capacity) return (uint64_t)-1;
    parallel_for(v->n_segments, n_threads, [&](uint64_t lo, uint64_t hi, int) {
        std::vector<int64_t> ts;
        std::vector<float> val;
        for (uint64_t i = lo; i < hi; i++) {
            ts.clear();
            val.clear();
            segment_grid(row_of(v, i), ts, val);
            std::memcpy(ts_out + point_off[i], ts.data(), ts.size() * 8);
            std::memcpy(val_out + point_off[i], val.data(), val.size() * 4);
        }
    });
    return total;
}

void mdbo_segment_sums(const mdbo_segments_view *v, float *sum_out, int n_threads) {
    parallel_for(v->n_segments, n_threads, [&](uint64_t lo, uint64_t hi, int) {
        for (uint64_t i = lo; i < hi; i++) sum_out[i] = segment_sum(row_of(v, i));
    });
}

void mdbo_aggregate(const mdbo_segments_view *v, const uint64_t *group_off, uint64_t n_groups, int64_t *count,
                    float *min, float *max, double *sum, int n_threads) {
    uint64_t whole[2] = {0, v->n_segments};
    if (!group_off) { group_off = whole; n_groups = 1; }
    parallel_for(n_groups, n_threads, [&](uint64_t lo, uint64_t hi, int) {
        for (uint64_t g = lo; g < hi; g++) {
            // model_simple_aggregates.rs:345-356 (count), :395-401 (min from f32::MAX),
            // :438-444 (max from f32::MIN), :481-511 (sum: f64 += f32 per-row sum)
            int64_t c = 0;
            float mn = std::numeric_limits<float>::max();
            float mx = std::numeric_limits<float>::lowest();
            double s = 0.0;
            for (uint64_t i = group_off[g]; i < group_off[g + 1]; i++) {
                Row r = row_of(v, i);
                c += (int64_t)segment_len(r.start_time, r.end_time, r.timestamps, r.n_timestamps);
                mn = rust_min(mn, r.min_value);
                mx = rust_max(mx, r.max_value);
                s += (double)segment_sum(r);
            }
            count[g] = c; min[g] = mn; max[g] = mx; sum[g] = s;
        }
    });
}

} // extern "C"
