// Host SIMT emulator for kernel-logic tests (TEST INFRASTRUCTURE ONLY — never part of the product).
//
// A CTA is run as `blockDim` cooperative fibers (ucontext) on ONE host thread.  A fiber runs until
// it reaches a collective (__syncthreads*, __ballot_sync, __any_sync, __shfl_xor_sync, ...),
// deposits its operand and yields; the last arriver completes the collective and everybody
// resumes.  All collectives must be called with the full mask by all 32 lanes of a warp (the
// kernels under test are written that way); a lane that skips one deadlocks the CTA, which the
// scheduler reports as an error instead of hanging.  `__shared__` becomes function-local `static`
// (one CTA runs at a time), atomics are plain read-modify-writes (one fiber runs at a time).
#pragma once
#include <ucontext.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
typedef void* cudaStream_t;
typedef int cudaError_t;
static inline cudaError_t cudaGetLastError() { return 0; }
#define cudaSuccess 0

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))

namespace ts_emu {

constexpr int kMaxThreads = 1024;
constexpr size_t kStack = 256 * 1024;

struct WarpSync {
    int arrived = 0, gen = 0;
    uint32_t in[32];
    uint32_t out[2][32];
};
struct BlockSync {
    int arrived = 0, gen = 0;
    int acc_and = 1;
    int out[2];
};

struct State {
    dim3 tid, bid, gdim, bdim;
    int nthreads = 0, cur = 0, alive = 0;
    ucontext_t sched;
    ucontext_t ctx[kMaxThreads];
    char* stacks = nullptr;
    bool done[kMaxThreads];
    WarpSync ws[kMaxThreads / 32];
    BlockSync bs;
    std::function<void()> body;
    long progress = 0;     // bumped whenever a collective completes or a fiber finishes
    int order = 0;         // 0 forward, 1 reverse, 2 random, 3 warpfirst, 4 warplast, 5/6 the same with lanes reversed
    unsigned long long rng = 12345;
    bool deadlock = false;
};
inline State& st() { static State s; return s; }

inline void yield_() {
    State& s = st();
    int me = s.cur;
    swapcontext(&s.ctx[me], &s.sched);
}

inline void fiber_entry() {
    State& s = st();
    s.body();
    s.done[s.cur] = true;
    s.alive--;
    s.progress++;
    // like the hardware, threads that have exited no longer take part in block barriers
    BlockSync& b = s.bs;
    if (b.arrived > 0 && b.arrived >= s.alive) {
        b.out[b.gen & 1] = b.acc_and;
        b.acc_and = 1;
        b.arrived = 0;
        b.gen++;
    }
    swapcontext(&s.ctx[s.cur], &s.sched);
}

// Runs `body` once per thread of every block of the grid.  Returns 0, or -1 on a deadlock
// (some lanes wait in a collective the others never reach).
inline int launch(dim3 grid, dim3 block, std::function<void()> body) {
    State& s = st();
    if (!s.stacks) {
        s.stacks = (char*)malloc(kStack * kMaxThreads);
        const char* o = getenv("TS_EMU_ORDER");
        s.order = !o ? 0 : !strcmp(o, "reverse") ? 1 : !strcmp(o, "random") ? 2 : !strcmp(o, "warpfirst") ? 3
                  : !strcmp(o, "warplast") ? 4 : !strcmp(o, "warpfirst-lanerev") ? 5
                  : !strcmp(o, "warplast-lanerev") ? 6 : 0;
    }
    const int nthreads = (int)(block.x * block.y * block.z);
    s.body = body;
    s.nthreads = nthreads;
    s.gdim = grid;
    s.bdim = block;
    for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
        for (unsigned bx = 0; bx < grid.x; ++bx) {
            s.bid = dim3(bx, by, bz);
            s.alive = nthreads;
            s.bs = BlockSync();
            for (int w = 0; w < (nthreads + 31) / 32; ++w) s.ws[w] = WarpSync();
            for (int t = 0; t < nthreads; ++t) {
                s.done[t] = false;
                getcontext(&s.ctx[t]);
                s.ctx[t].uc_stack.ss_sp = s.stacks + kStack * t;
                s.ctx[t].uc_stack.ss_size = kStack;
                s.ctx[t].uc_link = &s.sched;
                makecontext(&s.ctx[t], (void (*)())fiber_entry, 0);
            }
            auto resume = [&](int t) {
                s.cur = t;
                s.tid = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
                swapcontext(&s.sched, &s.ctx[t]);
            };
            // Scheduling order (TS_EMU_ORDER).  Fibers only yield at collectives, so a kernel that reads
            // shared memory another warp (or lane) writes WITHOUT a barrier / __syncwarp in between
            // can still see the right data under one order and stale data under another.
            //   forward / reverse / random : every round resumes every fiber once, in that order (warps
            //                                stay within one collective of each other);
            //   warpfirst / warplast       : one warp at a time runs as far as it can (until all of
            //                                its lanes wait at a block barrier or have exited), lowest
            //                                or highest warp first: the maximal skew between warps,
            //                                which is what exposes a missing __syncthreads;
            //   warpfirst-lanerev / warplast-lanerev : the same with the lanes of a warp resumed 31..0
            //                                (a missing __syncwarp between a lane's shared-memory
            //                                write and another lane's read).
            const int nwarps = (nthreads + 31) / 32;
            while (s.alive > 0) {
                long before = s.progress;
                if (s.order >= 3) {
                    for (int wi = 0; wi < nwarps; ++wi) {
                        const int w = (s.order & 1) ? wi : nwarps - 1 - wi;     // 3, 5: lowest warp first
                        for (;;) {      // run warp w while its own lanes keep completing collectives
                            long b2 = s.progress;
                            for (int li = 0; li < 32; ++li) {
                                const int l = s.order >= 5 ? 31 - li : li;          // 5, 6: lanes in reverse
                                const int t = w * 32 + l;
                                if (t < nthreads && !s.done[t]) resume(t);
                            }
                            if (s.progress == b2) break;
                        }
                    }
                } else {
                    // random order: a fresh permutation per round (start + slot * odd stride modulo the
                    // next power of two, skipping indices past the block)
                    unsigned pow2 = 1;
                    while ((int)pow2 < nthreads) pow2 <<= 1;
                    s.rng = s.rng * 6364136223846793005ull + 1442695040888963407ull;
                    const unsigned start = (unsigned)(s.rng >> 33) & (pow2 - 1), stride = ((unsigned)(s.rng >> 13) & (pow2 - 1)) | 1u;
                    for (int slot = 0; slot < (s.order == 2 ? (int)pow2 : nthreads); ++slot) {
                        int t = slot;
                        if (s.order == 1) t = nthreads - 1 - slot;
                        else if (s.order == 2) {
                            t = (int)((start + (unsigned)slot * stride) & (pow2 - 1));
                            if (t >= nthreads) continue;
                        }
                        if (s.done[t]) continue;
                        resume(t);
                    }
                }
                if (s.progress == before) {   // a full round without any collective completing
                    fprintf(stderr, "[ts_emu] deadlock in block (%u,%u): %d fibers stuck\n", bx, by, s.alive);
                    s.deadlock = true;
                    return -1;
                }
            }
        }
    return 0;
}

inline const uint32_t* warp_exchange(uint32_t v) {
    State& s = st();
    const int t = s.cur, lane = t & 31;
    WarpSync& w = s.ws[t >> 5];
    const int g = w.gen;
    w.in[lane] = v;
    if (++w.arrived == 32) {
        memcpy(w.out[g & 1], w.in, sizeof(w.in));
        w.arrived = 0;
        w.gen++;
        s.progress++;
    } else {
        while (w.gen == g) yield_();
    }
    return w.out[g & 1];
}

inline int block_barrier(int pred) {
    State& s = st();
    BlockSync& b = s.bs;
    const int g = b.gen;
    b.acc_and = b.acc_and && pred;
    if (++b.arrived >= s.alive) {
        b.out[g & 1] = b.acc_and;
        b.acc_and = 1;
        b.arrived = 0;
        b.gen++;
        s.progress++;
    } else {
        while (b.gen == g) yield_();
    }
    return b.out[g & 1];
}

inline int launch(dim3 grid, int nthreads, std::function<void()> body) {
    return launch(grid, dim3((unsigned)nthreads), body);
}

}  // namespace ts_emu

#define threadIdx (ts_emu::st().tid)
#define blockIdx (ts_emu::st().bid)
#define gridDim (ts_emu::st().gdim)

static inline void emu_require_full(unsigned mask) {
    if (mask != 0xffffffffu) { fprintf(stderr, "[ts_emu] partial-mask collective not supported\n"); abort(); }
}
static inline void __syncthreads() { ts_emu::block_barrier(1); }
static inline int __syncthreads_and(int p) { return ts_emu::block_barrier(p != 0); }
static inline void __syncwarp(unsigned mask = 0xffffffffu) { emu_require_full(mask); ts_emu::warp_exchange(0); }
static inline unsigned __ballot_sync(unsigned mask, int p) {
    emu_require_full(mask);
    const uint32_t* v = ts_emu::warp_exchange(p ? 1u : 0u);
    unsigned r = 0;
    for (int i = 0; i < 32; ++i) r |= (v[i] & 1u) << i;
    return r;
}
static inline int __any_sync(unsigned mask, int p) { return __ballot_sync(mask, p) != 0u; }
static inline int __all_sync(unsigned mask, int p) { return __ballot_sync(mask, p) == 0xffffffffu; }
static inline float __shfl_xor_sync(unsigned mask, float x, int d) {
    emu_require_full(mask);
    uint32_t u;
    memcpy(&u, &x, 4);
    const uint32_t* v = ts_emu::warp_exchange(u);
    uint32_t r = v[(ts_emu::st().cur & 31) ^ d];
    float f;
    memcpy(&f, &r, 4);
    return f;
}
static inline uint64_t __shfl_xor_sync(unsigned mask, uint64_t x, int d) {
    // two 32-bit exchanges, like the hardware
    emu_require_full(mask);
    const int lane = ts_emu::st().cur & 31;
    const uint32_t lo = ts_emu::warp_exchange((uint32_t)x)[lane ^ d];
    const uint32_t hi = ts_emu::warp_exchange((uint32_t)(x >> 32))[lane ^ d];
    return ((uint64_t)hi << 32) | lo;
}
static inline int __shfl_up_sync(unsigned mask, int x, int d) {
    emu_require_full(mask);
    const int lane = ts_emu::st().cur & 31;
    const uint32_t* v = ts_emu::warp_exchange((uint32_t)x);
    return lane >= d ? (int)v[lane - d] : x;
}
static inline void __threadfence() {}
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
// glibc declares (but does not export) __logf/__expf: map the CUDA fast-math names with macros
#define __logf(x) logf(x)
#define __expf(x) expf(x)
static inline int __shfl_xor_sync(unsigned mask, int x, int d) {
    emu_require_full(mask);
    const uint32_t* v = ts_emu::warp_exchange((uint32_t)x);
    return (int)v[(ts_emu::st().cur & 31) ^ d];
}
template <typename T>
static inline T emu_shfl_idx(T x, int src) {
    static_assert(sizeof(T) == 4, "32-bit shuffles only");
    uint32_t u;
    memcpy(&u, &x, 4);
    const uint32_t* v = ts_emu::warp_exchange(u);
    uint32_t r = v[src & 31];
    T out;
    memcpy(&out, &r, 4);
    return out;
}
static inline float __shfl_sync(unsigned mask, float x, int src) { emu_require_full(mask); return emu_shfl_idx(x, src); }
static inline int __shfl_sync(unsigned mask, int x, int src) { emu_require_full(mask); return emu_shfl_idx(x, src); }
static inline unsigned __shfl_sync(unsigned mask, unsigned x, int src) { emu_require_full(mask); return emu_shfl_idx(x, src); }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffs(unsigned x) { return __builtin_ffs((int)x); }
template <typename T>
static inline T __ldg(const T* p) { return *p; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
struct uchar4 { unsigned char x, y, z, w; };
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float atomicAdd(float* p, float v) { float o = *p; *p = o + v; return o; }
static inline int atomicAdd(int* p, int v) { int o = *p; *p = o + v; return o; }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { unsigned o = *p; *p = o + v; return o; }
static inline void atomicAdd(float2* p, float2 v) { p->x += v.x; p->y += v.y; }
static inline void atomicAdd(float4* p, float4 v) { p->x += v.x; p->y += v.y; p->z += v.z; p->w += v.w; }
static inline unsigned long long atomicCAS(unsigned long long* p, unsigned long long cmp, unsigned long long v) {
    unsigned long long o = *p;
    if (o == cmp) *p = v;
    return o;
}
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline int atomicMax(int* p, int v) { int o = *p; *p = std::max(o, v); return o; }
using std::max;
using std::min;
