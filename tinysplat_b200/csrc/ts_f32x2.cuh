// Packed fp32 pairs: Blackwell issues two IEEE fp32 operations per lane in one instruction
// (PTX add/sub/mul/fma.rn.f32x2 -> SASS FADD2 / FMUL2 / FFMA2 on sm_100a).  ptxas takes a broadcast
// scalar (`R.F32`) or a negated pair as an operand directly, so bcast() and neg() cost nothing.
// Each half rounds exactly like the scalar .rn instruction (no .ftz here, as the scalar intrinsics
// the kernels use elsewhere): a value computed through a pair is bit-identical to the scalar one —
// blend-backward relies on that to re-derive forward's skip decisions.
// Host emulator (TS_HOST_EMU, tests/emu): the same operations on two floats.
#pragma once
#include "ts_common.cuh"

namespace ts {

#ifdef TS_HOST_EMU
struct f32x2 { float lo, hi; };
__device__ __forceinline__ f32x2 pack2(float a, float b) { return f32x2{a, b}; }
__device__ __forceinline__ float lo2(f32x2 a) { return a.lo; }
__device__ __forceinline__ float hi2(f32x2 a) { return a.hi; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { return f32x2{a.lo + b.lo, a.hi + b.hi}; }
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) { return f32x2{a.lo - b.lo, a.hi - b.hi}; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { return f32x2{a.lo * b.lo, a.hi * b.hi}; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { return f32x2{fmaf(a.lo, b.lo, c.lo), fmaf(a.hi, b.hi, c.hi)}; }
#else
struct f32x2 { unsigned long long v; };
__device__ __forceinline__ f32x2 pack2(float a, float b) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ float lo2(f32x2 a) {
    float x, y;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a.v));
    return x;
}
__device__ __forceinline__ float hi2(f32x2 a) {
    float x, y;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a.v));
    return y;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
    return r;
}
#endif

__device__ __forceinline__ f32x2 bcast2(float a) { return pack2(a, a); }
__device__ __forceinline__ f32x2 neg2(f32x2 a) { return pack2(-lo2(a), -hi2(a)); }
__device__ __forceinline__ float hsum2(f32x2 a) { return lo2(a) + hi2(a); }

// eval_power (ts_blend_common.cuh) for two rows of one pixel column at once: the same operations in
// the same order, so each half equals the scalar result bit for bit.
//   dx = q0.x - px,  axd = q1.x * dx  (shared by the rows),  dy = q0.y - py
//   pw = fma(fma(q1.y, dy, axd), dx, (q1.z * dy) * dy)
__device__ __forceinline__ f32x2 eval_power2(float dx, float axd, float qy, float B, float C, f32x2 py, f32x2& dy) {
    dy = sub2(bcast2(qy), py);
    const f32x2 t = fma2(bcast2(B), dy, bcast2(axd));
    return fma2(t, bcast2(dx), mul2(mul2(bcast2(C), dy), dy));
}

}  // namespace ts
