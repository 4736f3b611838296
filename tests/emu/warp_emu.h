// warp_emu.h -- DEBUGGING HARNESS: runs warp-cooperative device code (modelardb_rs_b200/csrc/mdb_fit_warp.cuh) on the host.
//
// The 32 lanes of a warp are 32 cooperative fibers (ucontext) on one OS thread.  Every warp-level primitive
// (__shfl*_sync, __ballot_sync, __any_sync, __reduce_*_sync, __syncwarp) is a rendezvous: the calling lane publishes its
// operand and yields; the scheduler resumes the lanes round-robin, so when a lane continues, all 32 operands of that
// primitive are there.  Operands live in two alternating buffers, because a lane may already publish primitive k + 1
// while later lanes still read primitive k.  This only works for code whose lanes execute the same sequence of
// primitives -- which is exactly the property (uniform control flow) the device code relies on; a violation shows up
// as a deadlock-free but wrong exchange and is caught by `lanes_in_step`.
//
// Nothing under modelardb_rs_b200/ includes this file; the product is compiled by nvcc only.
#pragma once

#include <ucontext.h>

#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

namespace warp_emu {

constexpr int LANES = 32;

struct Warp {
    ucontext_t sched;
    ucontext_t ctx[LANES];
    std::vector<char> stack[LANES];
    uint64_t buf[2][LANES];
    uint32_t primitives[LANES]; // how many primitives each lane has issued
    bool done[LANES];
    int cur = 0;
    std::function<void(int)> body;
};

inline Warp *&current() {
    static Warp *w = nullptr;
    return w;
}

inline int lane() { return current()->cur; }

inline void yield() {
    Warp &w = *current();
    swapcontext(&w.ctx[w.cur], &w.sched);
}

// Publishes `v`, waits for the other lanes, returns the 32 operands of this primitive.
template <typename T> inline const uint64_t *rendezvous(T v) {
    static_assert(sizeof(T) <= 8, "operand size");
    Warp &w = *current();
    const int me = w.cur;
    const int half = (int)(w.primitives[me]++ & 1u);
    uint64_t raw = 0;
    std::memcpy(&raw, &v, sizeof(T));
    w.buf[half][me] = raw;
    yield();
    for (int l = 0; l < LANES; l++)
        if (!w.done[l] && w.primitives[l] != w.primitives[me] && w.primitives[l] != w.primitives[me] + 1 && w.primitives[l] + 1 != w.primitives[me]) {
            std::fprintf(stderr, "warp_emu: lanes are not in step (lane %d at %u, lane %d at %u)\n", me, w.primitives[me], l, w.primitives[l]);
            std::abort();
        }
    return w.buf[half];
}

template <typename T> inline T read_slot(const uint64_t *slots, int l) {
    T out;
    std::memcpy(&out, &slots[l & (LANES - 1)], sizeof(T));
    return out;
}

inline void trampoline() {
    Warp &w = *current();
    const int me = w.cur;
    w.body(me);
    w.done[me] = true;
    while (true) yield(); // never resumed again
}

// Runs body(lane) for the 32 lanes of one warp to completion.
inline void run(const std::function<void(int)> &body) {
    Warp w;
    w.body = body;
    current() = &w;
    for (int l = 0; l < LANES; l++) {
        w.stack[l].resize(512 * 1024);
        w.primitives[l] = 0;
        w.done[l] = false;
        getcontext(&w.ctx[l]);
        w.ctx[l].uc_stack.ss_sp = w.stack[l].data();
        w.ctx[l].uc_stack.ss_size = w.stack[l].size();
        w.ctx[l].uc_link = nullptr;
        makecontext(&w.ctx[l], trampoline, 0);
    }
    while (true) {
        bool any = false;
        for (int l = 0; l < LANES; l++) {
            if (w.done[l]) continue;
            any = true;
            w.cur = l;
            swapcontext(&w.sched, &w.ctx[l]);
        }
        if (!any) break;
    }
    current() = nullptr;
}

} // namespace warp_emu

// ---- the CUDA names the device code uses -----------------------------------------------------------------------------

struct EmuThreadIdx {
    struct X {
        operator unsigned() const { return (unsigned)warp_emu::lane(); }
    } x;
};
static EmuThreadIdx threadIdx;

template <typename T> static inline T __shfl_sync(unsigned, T v, int src) { return warp_emu::read_slot<T>(warp_emu::rendezvous(v), src); }
template <typename T> static inline T __shfl_up_sync(unsigned, T v, unsigned d) {
    const int me = warp_emu::lane();
    const uint64_t *slots = warp_emu::rendezvous(v);
    return me >= (int)d ? warp_emu::read_slot<T>(slots, me - (int)d) : v;
}
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int m) {
    const int me = warp_emu::lane();
    return warp_emu::read_slot<T>(warp_emu::rendezvous(v), me ^ m);
}
static inline unsigned __ballot_sync(unsigned, bool pred) {
    const uint64_t *slots = warp_emu::rendezvous<uint32_t>(pred ? 1u : 0u);
    unsigned m = 0;
    for (int l = 0; l < warp_emu::LANES; l++) m |= (warp_emu::read_slot<uint32_t>(slots, l) & 1u) << l;
    return m;
}
static inline bool __any_sync(unsigned mask, bool pred) { return __ballot_sync(mask, pred) != 0; }
static inline void __syncwarp(unsigned = 0xffffffffu) { warp_emu::rendezvous<uint32_t>(0u); }
static inline int __reduce_min_sync(unsigned, int v) {
    const uint64_t *slots = warp_emu::rendezvous(v);
    int m = INT_MAX;
    for (int l = 0; l < warp_emu::LANES; l++) m = std::min(m, warp_emu::read_slot<int>(slots, l));
    return m;
}
static inline int __reduce_max_sync(unsigned, int v) {
    const uint64_t *slots = warp_emu::rendezvous(v);
    int m = INT_MIN;
    for (int l = 0; l < warp_emu::LANES; l++) m = std::max(m, warp_emu::read_slot<int>(slots, l));
    return m;
}

static inline int __double2hiint(double d) { uint64_t u; std::memcpy(&u, &d, 8); return (int)(u >> 32); }
static inline double __hiloint2double(int hi, int lo) {
    uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
    double d;
    std::memcpy(&d, &u, 8);
    return d;
}
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
static inline double __longlong_as_double(long long v) { double d; std::memcpy(&d, &v, 8); return d; }
