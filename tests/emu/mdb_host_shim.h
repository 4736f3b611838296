// mdb_host_shim.h (tests/emu/; found through -I by the emulator build only) -- host stand-ins for the CUDA intrinsics used by the per-thread kernel bodies, so
// that tests/emu can step those bodies on the GPU-less build container (g++ -ffp-contract=off).
// NOT used by the product library: libmodelardb_cuda.so is compiled by nvcc and never sees this file.
#pragma once
#ifndef __CUDACC__
#include <cmath>
#include <cstdint>
#include <cstring>

static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __double2float_rn(double a) { return (float)a; }
static inline float __ull2float_rn(unsigned long long a) { return (float)a; }
static inline double __ull2double_rn(unsigned long long a) { return (double)a; }
static inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned)x); }
static inline int __clzll(long long x) { return x == 0 ? 64 : __builtin_clzll((unsigned long long)x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
using std::isinf;

// single-threaded equivalents of the inter-warp synchronisation words (mdb_device.cuh)
namespace mdb {
static inline uint32_t sync_cas(uint32_t *p, uint32_t expect, uint32_t val) { uint32_t old = *p; if (old == expect) *p = val; return old; }
static inline uint32_t sync_exch(uint32_t *p, uint32_t val) { uint32_t old = *p; *p = val; return old; }
static inline uint32_t sync_add(uint32_t *p, uint32_t val) { uint32_t old = *p; *p += val; return old; }
static inline uint32_t sync_load(const uint32_t *p) { return *p; }
static inline void sync_store(uint32_t *p, uint32_t v) { *p = v; }
static inline void sync_store8(uint8_t *p, uint8_t v) { *p = v; }
static inline void sync_fence() {}
static inline void sync_pause() {}
template <typename T> static inline T load_shared_record(const T *p) { return *p; }
} // namespace mdb
#endif
