// warp_emu.h -- DEBUGGING HARNESS: runs warp-cooperative device code (modelardb_rs_b200/csrc/mdb_fit_warp.cuh) on the host.
//
// The 32 lanes of a warp are 32 cooperative fibers on one OS thread (each with its own stack; switched by hand).  Every warp-level primitive
// (__shfl*_sync, __ballot_sync, __any_sync, __reduce_*_sync, __syncwarp) is a rendezvous: the calling lane publishes its
// operand and yields; the scheduler resumes the lanes round-robin, so when a lane continues, all 32 operands of that
// primitive are there.  Operands live in two alternating buffers, because a lane may already publish primitive k + 1
// while later lanes still read primitive k.  This only works for code whose lanes execute the same sequence of
// primitives -- which is exactly the property (uniform control flow) the device code relies on; a violation shows up
// as a deadlock-free but wrong exchange and is caught by `lanes_in_step`.
//
// Nothing under modelardb_rs_b200/ includes this file; the product is compiled by nvcc only.
#pragma once

#include <ucontext.h> // only used where the hand-written switch below is not available

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

// Switching between the lanes.  swapcontext() saves and restores the signal mask with two system calls per switch, which
// made the emulated tests the slowest part of the CPU suite; on x86-64 the switch is done by hand instead (callee-saved
// registers and the stack pointer, System V ABI).  Other architectures keep ucontext.
#if defined(__x86_64__)
#define WARP_EMU_FAST_SWITCH 1
extern "C" void warp_emu_switch(void **save_sp, void *load_sp);
asm(R"(
.text
.weak warp_emu_switch
.type warp_emu_switch,@function
warp_emu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size warp_emu_switch,.-warp_emu_switch
)");
#else
#define WARP_EMU_FAST_SWITCH 0
#endif

struct Warp {
#if WARP_EMU_FAST_SWITCH
    void *sched_sp = nullptr;
    void *sp[LANES];
#else
    ucontext_t sched;
    ucontext_t ctx[LANES];
#endif
    char *stack[LANES]; // from lane_stacks(): allocated once per host thread, reused by every run
    uint64_t buf[2][LANES];
    uint32_t primitives[LANES]; // how many primitives each lane has issued
    bool done[LANES];
    int cur = 0;
    std::function<void(int)> body;
};

inline Warp *&current() {
    static thread_local Warp *w = nullptr;
    return w;
}

inline int lane() { return current()->cur; }

inline void yield() {
    Warp &w = *current();
#if WARP_EMU_FAST_SWITCH
    warp_emu_switch(&w.sp[w.cur], w.sched_sp);
#else
    swapcontext(&w.ctx[w.cur], &w.sched);
#endif
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

#if WARP_EMU_FAST_SWITCH
extern "C" inline void warp_emu_trampoline() {
#else
inline void trampoline() {
#endif
    Warp &w = *current();
    const int me = w.cur;
    w.body(me);
    w.done[me] = true;
    while (true) yield(); // never resumed again
}

constexpr size_t STACK_BYTES = 512 * 1024;

// The lanes' stacks: one allocation per host thread for the life of the process (a fresh, zero-filled 16 MiB per run
// cost more in page faults than the emulated code in cycles).
inline char *lane_stacks() {
    static thread_local std::vector<char> stacks;
    if (stacks.empty()) stacks.resize(LANES * STACK_BYTES);
    return stacks.data();
}

// Runs body(lane) for the 32 lanes of one warp to completion.
inline void run(const std::function<void(int)> &body) {
    Warp w;
    w.body = body;
    current() = &w;
    for (int l = 0; l < LANES; l++) {
        w.stack[l] = lane_stacks() + (size_t)l * STACK_BYTES;
        w.primitives[l] = 0;
        w.done[l] = false;
#if WARP_EMU_FAST_SWITCH
        // a frame warp_emu_switch can "return" into: six callee-saved registers, the entry point, and a null return
        // address so that the entry point sees the stack alignment of a called function
        uintptr_t top = reinterpret_cast<uintptr_t>(w.stack[l] + STACK_BYTES) & ~uintptr_t(15);
        void **sp = reinterpret_cast<void **>(top);
        *--sp = nullptr;
        *--sp = reinterpret_cast<void *>(&warp_emu_trampoline);
        for (int r = 0; r < 6; r++) *--sp = nullptr;
        w.sp[l] = sp;
#else
        getcontext(&w.ctx[l]);
        w.ctx[l].uc_stack.ss_sp = w.stack[l];
        w.ctx[l].uc_stack.ss_size = STACK_BYTES;
        w.ctx[l].uc_link = nullptr;
        makecontext(&w.ctx[l], trampoline, 0);
#endif
    }
    while (true) {
        bool any = false;
        for (int l = 0; l < LANES; l++) {
            if (w.done[l]) continue;
            any = true;
            w.cur = l;
#if WARP_EMU_FAST_SWITCH
            warp_emu_switch(&w.sched_sp, w.sp[l]);
#else
            swapcontext(&w.sched, &w.ctx[l]);
#endif
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

static inline unsigned __reduce_min_sync(unsigned, unsigned v) {
    const uint64_t *slots = warp_emu::rendezvous(v);
    unsigned m = 0xffffffffu;
    for (int l = 0; l < warp_emu::LANES; l++) m = std::min(m, warp_emu::read_slot<unsigned>(slots, l));
    return m;
}
static inline unsigned __reduce_max_sync(unsigned, unsigned v) {
    const uint64_t *slots = warp_emu::rendezvous(v);
    unsigned m = 0u;
    for (int l = 0; l < warp_emu::LANES; l++) m = std::max(m, warp_emu::read_slot<unsigned>(slots, l));
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
