// K4, second generation — "row-pair" blend-forward (default; the first generation in blend.cu stays
// selectable with ts_set_blend_fwd_mode(0) / TS_BLEND_FWD=warp for A/B runs).
// Behind gsplat.rasterize_gaussians  [REF tinysplat/splatting/rasterize.py:44,50,83-86].
//
// Mapping.  One CTA of 128 threads per 16x16 tile [REF rasterize.py:19-20]; warp w owns the 8x8-pixel
// quadrant w of the tile; lane = (column 0..7, row pair 0..3) and composites TWO vertically adjacent
// pixels.  The two rows share dx, so the exponent, alpha, transmittance and colour arithmetic of both
// pixels is issued as packed fp32 pairs (FFMA2 / FMUL2 / FADD2, ts_f32x2.cuh): the first generation
// (one pixel per lane) ran at 88 % of the issue rate with the fp32 and XU pipes as its limiters, and
// the per-candidate overhead of the walk (bit scan, index, three LDS.128) is now paid once per two
// pixel rows instead of once per row.
// Each half of a pair rounds like the scalar instruction, in the same operation order as the first
// generation and as blend-backward: the images of the two generations are bit-identical, and
// backward re-derives exactly the skip decisions (pw < 0, alpha < 1/255, T <= 1e-4) taken here.
// Not HBM-bound: fp32 / MUFU issue.
#include "ts_blend_common.cuh"
#include "ts_f32x2.cuh"

namespace ts {

constexpr int kPThreads = 128;                 // 4 warps = the 4 quadrants of a tile
constexpr int kPBatch = 128;                   // records staged per batch: one per thread
constexpr int kPWords = kPBatch / 32;

// 4-bit mask: which 8x8 quadrants (bit 2*qy + qx) the footprint box of a staged record can reach.
__device__ __forceinline__ unsigned quadrant_mask(float4 q0) {
    const float X0 = (float)(blockIdx.x * kBlock) + kPixCenter;
    const float Y0 = (float)(blockIdx.y * kBlock) + kPixCenter;
    const float xl = q0.x - q0.z, xh = q0.x + q0.z, yl = q0.y - q0.w, yh = q0.y + q0.w;
    unsigned mx = 0u, my = 0u;
    if (xh >= X0 && xl <= X0 + 7.f) mx |= 1u;
    if (xh >= X0 + 8.f && xl <= X0 + 15.f) mx |= 2u;
    if (yh >= Y0 && yl <= Y0 + 7.f) my |= 1u;
    if (yh >= Y0 + 8.f && yl <= Y0 + 15.f) my |= 2u;
    return ((my & 1u) ? mx : 0u) | ((my & 2u) ? (mx << 2) : 0u);
}

template <int CH>
__global__ void __launch_bounds__(kPThreads)
blend_fwd_pair_kernel(int H, int W, int tbx, const int32_t* __restrict__ tile_offsets,
                      const int32_t* __restrict__ ids, const float4* __restrict__ recs,
                      const float* __restrict__ background, float* __restrict__ out_img,
                      float* __restrict__ out_ch3, float* __restrict__ final_T,
                      int32_t* __restrict__ n_contrib, int clamp_max1, int cap) {
    __shared__ __align__(16) float4 s_rec[2][kPBatch * 3];
    __shared__ unsigned s_mask[4][kPWords];     // [quadrant][staging warp]
    const unsigned full = 0xffffffffu;
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int j = blockIdx.x * kBlock + (warp & 1) * 8 + (lane & 7);
    const int i0 = blockIdx.y * kBlock + (warp >> 1) * 8 + 2 * (lane >> 3);
    const bool in_a = (i0 < H) && (j < W), in_b = (i0 + 1 < H) && (j < W);
    const float px = (float)j + kPixCenter;
    const f32x2 py = pack2((float)i0 + kPixCenter, (float)(i0 + 1) + kPixCenter);

    const int tile = blockIdx.y * tbx + blockIdx.x;
    const int start = __ldg(tile_offsets + tile);
    // cap = capacity of the id list: when the host sized it from an earlier step and this step needs
    // more (ts_bin_emit), a list that does not fit was neither filled nor sorted: skip the tile — the
    // host detects the overflow and renders again
    const int end = __ldg(tile_offsets + tile + 1);
    const int count = end <= cap ? end - start : 0;
    const int nb = (count + kPBatch - 1) / kPBatch;

    f32x2 T = pack2(1.f, 1.f);
    f32x2 acc[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) acc[c] = pack2(0.f, 0.f);
    int ncon_a = 0, ncon_b = 0;
    bool done_a = !in_a, done_b = !in_b;

    auto prefetch = [&](int b) {
        const int p = b * kPBatch + tid;
        if (p < count) {
            const int g = __ldg(ids + start + p);
            const float4* src = recs + 3 * (size_t)g;
            float4* dst = &s_rec[b & 1][tid * 3];
            cp_async16(dst, src);
            cp_async16(dst + 1, src + 1);
            cp_async16(dst + 2, src + 2);
        }
    };
    if (nb > 0) prefetch(0);
    cp_async_commit();

    for (int b = 0; b < nb; ++b) {
        const float4* rec = s_rec[b & 1];
        if (b + 1 < nb) prefetch(b + 1);
        cp_async_commit();
        cp_async_wait<1>();                     // batch b (this thread's copies) has landed
        unsigned mine = 0u;
        if (b * kPBatch + tid < count) mine = quadrant_mask(rec[tid * 3]);
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            const unsigned m = __ballot_sync(full, (mine >> s) & 1u);
            if (lane == 0) s_mask[s][warp] = m;
        }
        __syncthreads();                        // records + masks of batch b visible to all
        bool warp_done = __all_sync(full, done_a && done_b);
        for (int k = 0; k < kPWords && !warp_done; ++k) {
            unsigned m = s_mask[warp][k];
            while (m) {
                const int g = k * 32 + __ffs(m) - 1;
                m &= m - 1;
                const float4 q0 = rec[g * 3];
                const float4 q1 = rec[g * 3 + 1];
                const float dx = __fsub_rn(q0.x, px);
                f32x2 dy;
                const f32x2 pw = eval_power2(dx, __fmul_rn(q1.x, dx), q0.y, q1.y, q1.z, py, dy);
                const float pwa = lo2(pw), pwb = hi2(pw);
                const f32x2 araw = mul2(bcast2(q1.w), pack2(ex2_approx(-pwa), ex2_approx(-pwb)));
                const float ala = fminf(kAlphaMax, lo2(araw)), alb = fminf(kAlphaMax, hi2(araw));
                const bool oka = !done_a && pwa >= 0.f && ala >= kAlphaMin;
                const bool okb = !done_b && pwb >= 0.f && alb >= kAlphaMin;
                if (oka || okb) {
                    const f32x2 alpha = pack2(oka ? ala : 0.f, okb ? alb : 0.f);
                    const f32x2 nT = mul2(T, sub2(bcast2(1.f), alpha));       // = T where alpha is 0
                    const bool stop_a = oka && lo2(nT) <= kTStop, stop_b = okb && hi2(nT) <= kTStop;
                    done_a = done_a || stop_a;
                    done_b = done_b || stop_b;
                    // the Gaussian that would push T below the threshold does not contribute
                    const f32x2 wgt = mul2(pack2(stop_a ? 0.f : lo2(alpha), stop_b ? 0.f : hi2(alpha)), T);
                    const float4 q2 = rec[g * 3 + 2];
                    acc[0] = fma2(wgt, bcast2(q2.x), acc[0]);
                    if (CH > 1) acc[1] = fma2(wgt, bcast2(q2.y), acc[1]);
                    if (CH > 2) acc[2] = fma2(wgt, bcast2(q2.z), acc[2]);
                    if (CH > 3) acc[3] = fma2(wgt, bcast2(q2.w), acc[3]);
                    T = pack2(stop_a ? lo2(T) : lo2(nT), stop_b ? hi2(T) : hi2(nT));
                    const int idx = b * kPBatch + g + 1;
                    if (oka && !stop_a) ncon_a = idx;
                    if (okb && !stop_b) ncon_b = idx;
                }
            }
            warp_done = __all_sync(full, done_a && done_b);
        }
        // also guards reuse of s_rec[buf] / s_mask by the next iterations
        if (__syncthreads_and(warp_done)) break;
    }
    cp_async_wait<0>();

#pragma unroll
    for (int r = 0; r < 2; ++r) {
        if (!(r ? in_b : in_a)) continue;
        const float Tr = r ? hi2(T) : lo2(T);
        int ncon = r ? ncon_b : ncon_a;
        const size_t pix = (size_t)(i0 + r) * W + j;
        float a[CH];
#pragma unroll
        for (int c = 0; c < CH; ++c) a[c] = r ? hi2(acc[c]) : lo2(acc[c]);
        if (CH == 4 && out_ch3) {   // split output: RGB image + separate 4th-channel (depth) map
            // clamp_max1 folds the adapter's clamp(rgb, max=1) [REF rasterize.py:45] in; which
            // channels were clamped (zero gradient) is kept in the top bits of n_contrib
            unsigned cm = 0;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float o = fmaf(Tr, __ldg(background + c), a[c]);
                if (clamp_max1 && o > 1.f) { o = 1.f; cm |= 1u << c; }
                out_img[pix * 3 + c] = o;
            }
            out_ch3[pix] = fmaf(Tr, __ldg(background + 3), a[CH - 1]);
            ncon |= (int)(cm << kClampShift);
        } else {
#pragma unroll
            for (int c = 0; c < CH; ++c) out_img[pix * CH + c] = fmaf(Tr, __ldg(background + c), a[c]);
        }
        final_T[pix] = Tr;
        n_contrib[pix] = ncon;
    }
}

#ifndef TS_HOST_EMU
int launch_blend_fwd_pair(int CH, int H, int W, int tiles_x, int tiles_y, const int32_t* tile_offsets,
                          const int32_t* ids, const float* recs, const float* background, float* out_img,
                          float* out_ch3, float* final_T, int32_t* n_contrib, int clamp_max1, int cap,
                          cudaStream_t st) {
    dim3 grid(tiles_x, tiles_y);
#define TS_LAUNCH_FWD(C)                                                                                 \
    blend_fwd_pair_kernel<C><<<grid, kPThreads, 0, st>>>(H, W, tiles_x, tile_offsets, ids, (const float4*)recs, \
                                                         background, out_img, out_ch3, final_T, n_contrib,     \
                                                         clamp_max1, cap)
    switch (CH) {
        case 1: TS_LAUNCH_FWD(1); break;
        case 2: TS_LAUNCH_FWD(2); break;
        case 3: TS_LAUNCH_FWD(3); break;
        default: TS_LAUNCH_FWD(4); break;
    }
#undef TS_LAUNCH_FWD
    return 0;
}
#endif  // !TS_HOST_EMU

}  // namespace ts
