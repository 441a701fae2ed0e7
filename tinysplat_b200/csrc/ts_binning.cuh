// Tile-footprint helpers shared by the binning kernels and the fused project+count kernel.
#pragma once
#include "ts_common.cuh"

namespace ts {

constexpr int kCoopThreshold = 8;    // rects larger than this are expanded by the whole warp
// Per-tile counters / cursors live one per 128-byte line: L2 serialises atomics per line, and
// adjacent 4-byte counters (32 tiles per line) made the count/emit passes contention-bound.
constexpr int kCounterStride = 32;   // in int32 elements

// Tile rectangle [lo,hi) of a packed record: 3-sigma bbox ∩ opacity-aware footprint.
__device__ __forceinline__ void tile_rect(float4 q0, float radius, int tbx, int tby, int cull,
                                          int& lox, int& loy, int& hix, int& hiy) {
    tile_bbox(q0.x, q0.y, radius, tbx, tby, lox, loy, hix, hiy);
    if (cull) {
        // pixel centres of tile t along x: 16t + 0.5 .. 16t + 15.5
        float fl = ceilf((q0.x - q0.z - 15.5f) * (1.f / kBlock));
        float fh = floorf((q0.x + q0.z - 0.5f) * (1.f / kBlock));
        float gl = ceilf((q0.y - q0.w - 15.5f) * (1.f / kBlock));
        float gh = floorf((q0.y + q0.w - 0.5f) * (1.f / kBlock));
        fl = fminf(fmaxf(fl, -1.f), 1e9f); gl = fminf(fmaxf(gl, -1.f), 1e9f);
        fh = fminf(fmaxf(fh, -2.f), 1e9f); gh = fminf(fmaxf(gh, -2.f), 1e9f);
        lox = max(lox, (int)fl); loy = max(loy, (int)gl);
        hix = min(hix, (int)fh + 1); hiy = min(hiy, (int)gh + 1);
        if (hix < lox) hix = lox;
        if (hiy < loy) hiy = loy;
    }
}

// Run f(tile_id, payload) for every tile of every lane's rectangle.  Small rectangles are
// walked by their own lane; large ones are expanded cooperatively by the whole warp (payload
// broadcast from the owning lane) so that one screen-filling Gaussian does not serialise
// thousands of atomics on a single lane.  Must be reached by all 32 lanes.
template <typename F>
__device__ __forceinline__ void for_each_tile(int lox, int loy, int hix, int hiy, int tbx,
                                              uint32_t pay_lo, uint32_t pay_hi, F f) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    int w = hix - lox, h = hiy - loy;
    int n = (w > 0 && h > 0) ? w * h : 0;
    if (n > 0 && n <= kCoopThreshold) {
        for (int y = loy; y < hiy; ++y)
            for (int x = lox; x < hix; ++x) f(y * tbx + x, pay_lo, pay_hi);
    }
    __syncwarp(full);
    unsigned big = __ballot_sync(full, n > kCoopThreshold);
    while (big) {
        int src = __ffs(big) - 1;
        big &= big - 1;
        int slox = __shfl_sync(full, lox, src), sloy = __shfl_sync(full, loy, src);
        int sw = __shfl_sync(full, w, src), sn = __shfl_sync(full, n, src);
        uint32_t plo = __shfl_sync(full, pay_lo, src), phi = __shfl_sync(full, pay_hi, src);
        for (int k = lane; k < sn; k += 32) {
            int y = k / sw, x = k - y * sw;
            f((sloy + y) * tbx + slox + x, plo, phi);
        }
    }
}


// Half extents (pixels) of the axis-aligned box around {alpha >= 1/255}: sigma <= tau =
// ln(255*opac); the ellipse's box is sqrt(2*tau*cov_xx) x sqrt(2*tau*cov_yy), cov = conic^-1.
// cull == 0 disables culling (infinite extent); an opacity below 1/255 can never contribute.
__device__ __forceinline__ void footprint_extent(float a, float b, float c, float op, int cull,
                                                 float& hx, float& hy) {
    hx = 1e30f; hy = 1e30f;
    if (!cull) return;
    float det = a * c - b * b;
    if (!(op * 255.f >= 1.f)) { hx = -1e30f; hy = -1e30f; }
    else if (det > 0.f && a > 0.f && c > 0.f) {
        float two_tau = 2.f * (__logf(255.f * op) + 0.01f);
        hx = sqrtf(two_tau * c / det) * 1.001f + 0.01f;
        hy = sqrtf(two_tau * a / det) * 1.001f + 0.01f;
    }
}

}  // namespace ts
