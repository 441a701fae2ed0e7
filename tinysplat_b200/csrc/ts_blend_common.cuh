// Pieces shared by the blend kernels (blend.cu: forward and first-generation backward, one warp per
// 8x4 sub-block; blend_group.cu: grouped backward, one 8-lane group per sub-block, 4 rows per lane).
#pragma once
#include "ts_common.cuh"

namespace ts {

// Which tile a CTA works on.  The blend kernels run on a 1-D grid of T CTAs; `order` (ts_bin_tile_order:
// tiles by descending list length) makes the CTAs that start first take the longest lists, so the
// grid's tail consists of the cheapest tiles instead of whichever tiles the raster order put last.
// order == nullptr: CTA i works on tile i.
struct TileId { int tile, bx, by; };
__device__ __forceinline__ TileId tile_id(const int32_t* __restrict__ order, int tbx) {
    TileId t;
    t.tile = order ? __ldg(order + blockIdx.x) : (int)blockIdx.x;
    t.by = t.tile / tbx;
    t.bx = t.tile - t.by * tbx;
    return t;
}

constexpr int kClampShift = 28;                      // n_contrib bits 28..30: clamped-channel mask
constexpr int kCountMask = (1 << kClampShift) - 1;

// Exponent (in log2 units) of one Gaussian at one pixel.  Written with explicit fma/mul
// intrinsics so that forward and backward (of either kernel generation) round identically:
// backward must re-derive exactly the skip decisions (pw < 0, alpha < 1/255) forward took.
__device__ __forceinline__ float eval_power(const float4& q0, const float4& q1, float px, float py,
                                            float& dx, float& dy) {
    dx = __fsub_rn(q0.x, px);
    dy = __fsub_rn(q0.y, py);
    float t = __fmaf_rn(q1.y, dy, __fmul_rn(q1.x, dx));
    return __fmaf_rn(t, dx, __fmul_rn(__fmul_rn(q1.z, dy), dy));
}

// ---- exact per-row footprint of one packed record inside one 16x16 tile ----------------------
// Returns a 32-bit mask, bit (2*row + half): can any pixel of tile row `row` (0..15), columns
// 8*half .. 8*half+7, reach alpha >= 1/255?  Conservative by construction (a superset of the
// pixels the blend loop would accept), so using it to skip work cannot change the image:
//   accept  <=>  opac * 2^-pw >= 1/255  <=>  pw <= log2(255*opac) =: thr,
//   pw = A dx^2 + B dx dy + C dy^2 = A (dx + s dy)^2 + D dy^2,   s = B/2A,  D = C - B^2/4A,
// so on the row at distance dy the accepted columns are |dx + s dy| <= sqrt((thr - D dy^2)/A).
// Margins: thr gets +0.02 (ex2.approx / log2 error) plus the fp32 evaluation noise of pw at the
// far corner of the tile (2^-20 x the magnitude of its terms: far pixels of huge needle-shaped
// Gaussians are accepted or rejected by rounding noise and must stay in the superset); the
// column interval gets 0.02 px + 2^-19 of its magnitude.
// X0, Y0: pixel-centre coordinates of the tile's first pixel.  q0.z > 1e29 (culling disabled
// by the caller) -> all bits; q0.z < 0 (opacity below 1/255) -> none.
__host__ __device__ __forceinline__ unsigned footprint_rowmask(const float4 q0, const float4 q1,
                                                               float X0, float Y0) {
    if (q0.z > 1e29f) return 0xffffffffu;
    if (q0.z < 0.f) return 0u;
    const float A = q1.x, B = q1.y, C = q1.z;
    // a conic that is not positive along x (numerically broken covariance, NaN): no culling — the
    // blend loop's own pw >= 0 / alpha >= 1/255 tests decide, exactly as without the mask
    if (!(A > 0.f)) return 0xffffffffu;
    const float xr = q0.x - X0, yr = q0.y - Y0;       // centre relative to the tile's first pixel
    const float dxm = fmaxf(fabsf(xr), fabsf(xr - 15.f));
    const float dym = fmaxf(fabsf(yr), fabsf(yr - 15.f));
    const float noise = 9.5367431640625e-7f * (fabsf(A) * dxm * dxm + fabsf(B) * dxm * dym + fabsf(C) * dym * dym);
#ifdef __CUDA_ARCH__
    const float thr = __log2f(255.f * q1.w) + 0.02f + noise;
#else
    const float thr = log2f(255.f * q1.w) + 0.02f + noise;
#endif
    const float invA = 1.f / A;
    const float s = 0.5f * B * invA;
    const float D = C - 0.5f * B * s;
    // column margin: 0.03 px + rounding of q0.x - px (|x| 2^-23) and of the centre line s*dy
    const float mrg = 0.03f + 2e-6f * (fabsf(q0.x) + fabsf(X0) + fabsf(s) * dym);
    unsigned mask = 0u;
#pragma unroll
    for (int row = 0; row < 16; ++row) {
        const float dy = yr - (float)row;
        const float rem = thr - (D * dy) * dy;        // A (dx + s dy)^2 <= rem
#ifdef __CUDA_ARCH__
        float w;
        asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(w) : "f"(fmaxf(rem * invA, 0.f)));
#else
        const float w = sqrtf(fmaxf(rem * invA, 0.f));
#endif
        const float wm = w * 1.0002f + mrg;
        const float pc = xr + s * dy;                 // column (relative) of the row's centre line
        const float lo = pc - wm, hi = pc + wm;
        // every comparison is written so that a NaN keeps the bit (conservative)
        const bool miss = (rem < 0.f);
        const bool miss0 = miss || (hi < 0.f) || (lo > 7.f);
        const bool miss1 = miss || (hi < 8.f) || (lo > 15.f);
        mask |= (miss0 ? 0u : 1u) << (2 * row);
        mask |= (miss1 ? 0u : 2u) << (2 * row);
    }
    return mask;
}

}  // namespace ts
