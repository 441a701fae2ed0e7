// K4 / K5 — per-pixel front-to-back alpha compositing, forward and backward.
// Behind gsplat.rasterize_gaussians  [REF tinysplat/splatting/rasterize.py:44,50,83-86].
//
// One CTA per 16x16 tile [REF rasterize.py:19-20], 8 warps, each warp owns an 8x4-pixel
// sub-block.  The tile's depth-sorted id list is consumed in batches of 256: every thread
// gathers one 48-byte packed record (3 x LDGSTS.128, double buffered, L2-resident record
// array) into shared memory, then tests its Gaussian's alpha>=1/255 footprint against the 8
// sub-blocks; warp ballots turn that into one 32-bit mask per (sub-block, staging warp).  Each
// warp then walks only the set bits of ITS masks, so Gaussians that cannot touch its 32 pixels
// cost it nothing (hierarchical culling inside shared memory — no extra global traffic).
//
// Not HBM-bound: fp32 FMA/MUFU issue bound (forward at ~89 % of peak issue rate); backward adds a
// per-warp shared-memory reduction and L2 vector atomics (~72 % of peak issue rate).
#include <climits>
#include <cstdlib>
#include <cstring>
#include "ts_blend_common.cuh"

#ifndef TS_DEFAULT_BLEND_MODE
#define TS_DEFAULT_BLEND_MODE 1
#endif

namespace ts {

constexpr int kBlendThreads = 256;
constexpr int kBatch = 256;

// Geometry of the thread->pixel map shared by forward and backward.
struct PixMap {
    int warp, lane, i, j;
    bool inside;
    float px, py;
};

__device__ __forceinline__ PixMap pix_map(int H, int W, int bx, int by) {
    PixMap m;
    m.warp = threadIdx.x >> 5;
    m.lane = threadIdx.x & 31;
    int wx = m.warp & 1, wy = m.warp >> 1, lx = m.lane & 7, ly = m.lane >> 3;
    m.j = bx * kBlock + wx * 8 + lx;
    m.i = by * kBlock + wy * 4 + ly;
    m.inside = (m.i < H) && (m.j < W);
    m.px = (float)m.j + kPixCenter;
    m.py = (float)m.i + kPixCenter;
    return m;
}

// 8-bit mask: which of the tile's 8 sub-blocks (8 wide x 4 tall) the footprint box can reach.
__device__ __forceinline__ unsigned subblock_mask(float4 q0, int bx, int by) {
    const float X0 = (float)(bx * kBlock) + kPixCenter;
    const float Y0 = (float)(by * kBlock) + kPixCenter;
    float xl = q0.x - q0.z, xh = q0.x + q0.z, yl = q0.y - q0.w, yh = q0.y + q0.w;
    unsigned mx = 0, my = 0;
    // columns: sub-block wx covers pixel centres [X0+8wx, X0+8wx+7]
    if (xh >= X0 && xl <= X0 + 7.f) mx |= 1u;
    if (xh >= X0 + 8.f && xl <= X0 + 15.f) mx |= 2u;
    // rows: sub-block wy covers pixel centres [Y0+4wy, Y0+4wy+3]
#pragma unroll
    for (int wy = 0; wy < 4; ++wy)
        if (yh >= Y0 + 4.f * wy && yl <= Y0 + 4.f * wy + 3.f) my |= 1u << wy;
    unsigned m = 0;
#pragma unroll
    for (int wy = 0; wy < 4; ++wy)
        if (my & (1u << wy)) m |= mx << (2 * wy);
    return m;  // bit (2*wy + wx) == warp index of that sub-block
}

template <int CH>
__global__ void __launch_bounds__(kBlendThreads)
blend_fwd_kernel(int H, int W, int tbx, const int32_t* __restrict__ tile_offsets,
                 const int32_t* __restrict__ ids, const float4* __restrict__ recs,
                 const float* __restrict__ background, float* __restrict__ out_img,
                 float* __restrict__ out_ch3, float* __restrict__ final_T,
                 int32_t* __restrict__ n_contrib, int clamp_max1, int cap,
                 const int32_t* __restrict__ order) {
    __shared__ __align__(16) float4 s_rec[2][kBatch * 3];
    __shared__ unsigned s_mask[8][8];  // [sub-block][staging warp]
    // [sub-block][i]: staged index of the sub-block's i-th candidate.  The warp turns its eight mask words
    // into this byte list once per batch (one POPC pair per word) and walks it, instead of BREV + FLO +
    // two ALU operations per candidate (the XU pipe — MUFU.EX2, BREV, FLO, POPC — was at 46 %): -4 %
    __shared__ unsigned char s_list[8][kBatch];
    const unsigned full = 0xffffffffu;
    const TileId tl = tile_id(order, tbx);
    const PixMap pm = pix_map(H, W, tl.bx, tl.by);
    const int tid = threadIdx.x;
    const int tile = tl.tile;
    const int start = __ldg(tile_offsets + tile);
    // cap = capacity of the id list: when the host sized it from an earlier step and this step needs
    // more (ts_bin_emit), a list that does not fit was neither filled nor sorted: skip the tile — the
    // host detects the overflow and renders again
    const int end = __ldg(tile_offsets + tile + 1);
    const int count = end <= cap ? end - start : 0;
    const int nb = (count + kBatch - 1) / kBatch;

    float T = 1.f;
    float acc[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) acc[c] = 0.f;
    int ncon = 0;
    bool done = !pm.inside;

    auto prefetch = [&](int b) {
        int p = b * kBatch + tid;
        if (p < count) {
            int g = __ldg(ids + start + p);
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
        const int buf = b & 1;
        if (b + 1 < nb) prefetch(b + 1);
        cp_async_commit();
        cp_async_wait<1>();  // batch b (this thread's copies) has landed
        unsigned mine = 0;
        if (b * kBatch + tid < count) mine = subblock_mask(s_rec[buf][tid * 3], tl.bx, tl.by);
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            unsigned m = __ballot_sync(full, (mine >> s) & 1u);
            if (pm.lane == 0) s_mask[s][pm.warp] = m;
        }
        __syncthreads();  // records + masks of batch b visible to all
        bool warp_done = __all_sync(full, done);
        if (!warp_done) {
            int nlist = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const unsigned word = s_mask[pm.warp][k];
                if ((word >> pm.lane) & 1u)
                    s_list[pm.warp][nlist + __popc(word & ((1u << pm.lane) - 1u))] = (unsigned char)(k * 32 + pm.lane);
                nlist += __popc(word);
            }
            __syncwarp(full);
            for (int base = 0; base < nlist && !warp_done; base += 32) {
              const int end = min(nlist, base + 32);
              for (int i = base; i < end; ++i) {
                const int g = s_list[pm.warp][i];
                const float4 q0 = s_rec[buf][g * 3];
                const float4 q1 = s_rec[buf][g * 3 + 1];
                float dx, dy;
                float pw = eval_power(q0, q1, pm.px, pm.py, dx, dy);
                float alpha = fminf(kAlphaMax, __fmul_rn(q1.w, ex2_approx(-pw)));
                if (!done && pw >= 0.f && alpha >= kAlphaMin) {
                    float nT = T * (1.f - alpha);
                    if (nT <= kTStop) {
                        done = true;
                    } else {
                        const float4 q2 = s_rec[buf][g * 3 + 2];
                        float wgt = alpha * T;
                        acc[0] = fmaf(wgt, q2.x, acc[0]);
                        if (CH > 1) acc[1] = fmaf(wgt, q2.y, acc[1]);
                        if (CH > 2) acc[2] = fmaf(wgt, q2.z, acc[2]);
                        if (CH > 3) acc[3] = fmaf(wgt, q2.w, acc[3]);
                        T = nT;
                        ncon = b * kBatch + g + 1;
                    }
                }
              }
              warp_done = __all_sync(full, done);
            }
        }
        // also guards reuse of s_rec[buf] / s_mask by the next iterations
        if (__syncthreads_and(warp_done)) break;
    }
    cp_async_wait<0>();

    if (pm.inside) {
        const size_t pix = (size_t)pm.i * W + pm.j;
        if (CH == 4 && out_ch3) {   // split output: RGB image + separate 4th-channel (depth) map
            // clamp_max1 folds the adapter's clamp(rgb, max=1) [REF rasterize.py:45] in; which
            // channels were clamped (zero gradient) is kept in the top bits of n_contrib
            unsigned cm = 0;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float o = fmaf(T, __ldg(background + c), acc[c]);
                if (clamp_max1 && o > 1.f) { o = 1.f; cm |= 1u << c; }
                out_img[pix * 3 + c] = o;
            }
            out_ch3[pix] = fmaf(T, __ldg(background + 3), acc[CH - 1]);
            ncon |= (int)(cm << kClampShift);
        } else {
#pragma unroll
            for (int c = 0; c < CH; ++c) out_img[pix * CH + c] = fmaf(T, __ldg(background + c), acc[c]);
        }
        final_T[pix] = T;
        n_contrib[pix] = ncon;
    }
}

// Shared-memory plan of blend_bwd (dynamic, ~70 KB -> 3 CTAs/SM):
//   s_rec  [2][256*3] float4   double-buffered gathered records
//   s_acc  [256*12]   float    per-batch CTA accumulators (packed gradient layout)
//   s_part [8 warps][3 candidates][10 values][32 lanes] float   per-warp reduction staging
//   s_gid  [2][256] int, s_mask [8][8] unsigned, s_nmax
constexpr int kGroup = 3;                 // candidates reduced together (3*10 = 30 busy lanes)
constexpr int kPartFloats = 10 * 32;      // one candidate's staged partials
constexpr size_t kBwdSmemBytes = sizeof(float4) * 2 * kBatch * 3 + sizeof(float) * kBatch * kGradFloats +
                                 sizeof(float) * 8 * kGroup * kPartFloats + sizeof(int) * 2 * kBatch +
                                 sizeof(unsigned) * 64 + sizeof(int) * 8 * 4 + 16;

// GCH = number of colour channels that carry a cotangent (GCH <= CH; the fused RGB+depth pass
// with no depth loss has CH = 4, GCH = 3 and skips all channel-3 gradient arithmetic).
template <int CH, int GCH>
__global__ void __launch_bounds__(kBlendThreads)
blend_bwd_kernel(int H, int W, int tbx, const int32_t* __restrict__ tile_offsets,
                 const int32_t* __restrict__ ids, const float4* __restrict__ recs,
                 const float* __restrict__ background, const float* __restrict__ final_T,
                 const int32_t* __restrict__ n_contrib, const float* __restrict__ v_out_img,
                 const float* __restrict__ v_out_ch3, int split_ch3,
                 const float* __restrict__ v_out_alpha, float4* __restrict__ grads,
                 const int32_t* __restrict__ order) {
    TS_DYN_SMEM(unsigned char, s_raw, 16);
    float4* s_rec = reinterpret_cast<float4*>(s_raw);                         // [2][kBatch*3]
    float* s_acc = reinterpret_cast<float*>(s_rec + 2 * kBatch * 3);          // [kBatch*12]
    float* s_part = s_acc + kBatch * kGradFloats;                             // [8][kGroup][320]
    int* s_gid = reinterpret_cast<int*>(s_part + 8 * kGroup * kPartFloats);   // [2][kBatch]
    unsigned* s_mask = reinterpret_cast<unsigned*>(s_gid + 2 * kBatch);       // [8][8]
    int* s_slot = reinterpret_cast<int*>(s_mask + 64);                        // [8 warps][4]
    int* s_nmax = s_slot + 32;
    const unsigned full = 0xffffffffu;
    const TileId tl = tile_id(order, tbx);
    const PixMap pm = pix_map(H, W, tl.bx, tl.by);
    const int tid = threadIdx.x;
    const int tile = tl.tile;
    const int start = __ldg(tile_offsets + tile);

    float T_final = 1.f, v_oa = 0.f;
    float v_out[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) v_out[c] = 0.f;
    int nc = 0;
    if (pm.inside) {
        const size_t pix = (size_t)pm.i * W + pm.j;
        T_final = __ldg(final_T + pix);
        nc = __ldg(n_contrib + pix);
        const unsigned cm = (unsigned)nc >> kClampShift;    // channels clamped by forward
        nc &= kCountMask;
        if (CH == 4 && split_ch3) {
#pragma unroll
            for (int c = 0; c < 3; ++c)
                v_out[c] = (v_out_img && !((cm >> c) & 1u)) ? __ldg(v_out_img + pix * 3 + c) : 0.f;
            if (GCH == 4) v_out[CH - 1] = v_out_ch3 ? __ldg(v_out_ch3 + pix) : 0.f;
        } else {
#pragma unroll
            for (int c = 0; c < CH; ++c) v_out[c] = __ldg(v_out_img + pix * CH + c);
        }
        if (v_out_alpha) v_oa = __ldg(v_out_alpha + pix);
    }
    float bgdot = 0.f;
#pragma unroll
    for (int c = 0; c < GCH; ++c) bgdot = fmaf(__ldg(background + c), v_out[c], bgdot);
    const float wfin = T_final * (v_oa - bgdot);
    float T = T_final;
    float buffer[GCH];
#pragma unroll
    for (int c = 0; c < GCH; ++c) buffer[c] = 0.f;

    if (tid == 0) *s_nmax = 0;
#pragma unroll
    for (int k = 0; k < kGradFloats; ++k) s_acc[tid * kGradFloats + k] = 0.f;
    __syncthreads();
    int wmax = nc;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) wmax = max(wmax, __shfl_xor_sync(full, wmax, d));
    if (pm.lane == 0) atomicMax(s_nmax, wmax);
    __syncthreads();
    const int nmax = *s_nmax;  // entries [0, nmax) of the tile list contributed somewhere
    const int nb = (nmax + kBatch - 1) / kBatch;

    // this lane's role in the group reduction: row (candidate rc, value rv) of s_part
    float* my_part = s_part + pm.warp * kGroup * kPartFloats;
    const int rc = pm.lane / 10, rv = pm.lane - 10 * rc;
    const int racc = rv < 6 ? rv : rv + 2;    // slot -> float offset in the packed gradient record

    // batch b, slot t  <->  list position  p = nmax-1 - (b*256 + t)   (back to front)
    auto prefetch = [&](int b) {
        int p = nmax - 1 - (b * kBatch + tid);
        if (p >= 0) {
            int g = __ldg(ids + start + p);
            s_gid[(b & 1) * kBatch + tid] = g;
            const float4* src = recs + 3 * (size_t)g;
            float4* dst = s_rec + (b & 1) * kBatch * 3 + tid * 3;
            cp_async16(dst, src);
            cp_async16(dst + 1, src + 1);
            cp_async16(dst + 2, src + 2);
        }
    };
    if (nb > 0) prefetch(0);
    cp_async_commit();

    for (int b = 0; b < nb; ++b) {
        const int buf = b & 1;
        const float4* rec = s_rec + buf * kBatch * 3;
        if (b + 1 < nb) prefetch(b + 1);
        cp_async_commit();
        cp_async_wait<1>();
        const int pbase = nmax - 1 - b * kBatch;
        const int my_p = pbase - tid;
        unsigned mine = 0;
        if (my_p >= 0) mine = subblock_mask(rec[tid * 3], tl.bx, tl.by);
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            unsigned m = __ballot_sync(full, (mine >> s) & 1u);
            if (pm.lane == 0) s_mask[s * 8 + pm.warp] = m;
        }
        __syncthreads();

        int gcount = 0;                           // candidates staged in the current group
        int* my_slot = s_slot + pm.warp * 4;      // their batch slots
        // sums the staged rows: 30 lanes x (8 LDS.128 + 32 FADD), then one shared atomic per lane
        auto flush_group = [&]() {
            __syncwarp(full);
            if (pm.lane < gcount * 10) {
                const float4* row = reinterpret_cast<const float4*>(my_part + (rc * 10 + rv) * 32);
                float4 acc4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float4 t = row[(j + pm.lane) & 7];   // rotated start: conflict-free LDS.128
                    acc4.x += t.x; acc4.y += t.y; acc4.z += t.z; acc4.w += t.w;
                }
                atomicAdd(s_acc + my_slot[rc] * kGradFloats + racc, (acc4.x + acc4.y) + (acc4.z + acc4.w));
            }
            __syncwarp(full);
            gcount = 0;
        };

        for (int k = 0; k < 8; ++k) {
            unsigned m = s_mask[pm.warp * 8 + k];
            while (m) {
                int bit = __ffs(m) - 1;
                m &= m - 1;
                const int g = k * 32 + bit;
                const int p = pbase - g;
                const float4 q0 = rec[g * 3];
                const float4 q1 = rec[g * 3 + 1];
                float dx, dy;
                float pw = eval_power(q0, q1, pm.px, pm.py, dx, dy);
                float vis = ex2_approx(-pw);
                float araw = __fmul_rn(q1.w, vis);
                float alpha = fminf(kAlphaMax, araw);
                bool valid = pm.inside && (p < nc) && (pw >= 0.f) && (alpha >= kAlphaMin);
                if (!__any_sync(full, valid)) continue;
                float val[10];
#pragma unroll
                for (int v = 0; v < 10; ++v) val[v] = 0.f;
                if (valid) {
                    const float4 q2 = rec[g * 3 + 2];
                    const float col[4] = {q2.x, q2.y, q2.z, q2.w};
                    float ra = rcp_approx(1.f - alpha);
                    T *= ra;
                    float fac = alpha * T;
                    float v_alpha = wfin * ra;
#pragma unroll
                    for (int c = 0; c < GCH; ++c) {
                        val[6 + c] = fac * v_out[c];
                        v_alpha = fmaf(col[c] * T - buffer[c] * ra, v_out[c], v_alpha);
                        buffer[c] = fmaf(col[c], fac, buffer[c]);
                    }
                    if (araw > kAlphaMax) v_alpha = 0.f;  // clamped: d alpha / d araw = 0
                    float v_sig = -araw * v_alpha;
                    val[0] = v_sig * dx;
                    val[1] = v_sig * dy;
                    val[2] = val[0] * dx;
                    val[3] = val[0] * dy;
                    val[4] = val[1] * dy;
                    val[5] = vis * v_alpha;
                }
                float* pp = my_part + gcount * kPartFloats + pm.lane;
#pragma unroll
                for (int v = 0; v < 10; ++v) pp[v * 32] = (v < 6 + GCH) ? val[v] : 0.f;
                my_slot[gcount] = g;   // same value from every lane: one broadcast store
                if (++gcount == kGroup) flush_group();
            }
        }
        if (gcount) flush_group();
        __syncthreads();  // all warps finished batch b: s_acc complete
        if (mine) {
            float4* a4 = reinterpret_cast<float4*>(s_acc + tid * kGradFloats);
            float4* dst = grads + 3 * (size_t)s_gid[buf * kBatch + tid];
            atomicAdd(dst, a4[0]);
            atomicAdd(dst + 1, a4[1]);
            atomicAdd(dst + 2, a4[2]);
            a4[0] = make_float4(0.f, 0.f, 0.f, 0.f);
            a4[1] = make_float4(0.f, 0.f, 0.f, 0.f);
            a4[2] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        __syncthreads();  // s_acc reset + buffers free before the next batch touches them
    }
    cp_async_wait<0>();
}

template <int CH>
__global__ void __launch_bounds__(256)
unpack_grads_kernel(int N, const int32_t* __restrict__ radii, const float* __restrict__ conics,
                    const float4* __restrict__ grads, float2* __restrict__ v_xys,
                    float* __restrict__ v_conics, float* __restrict__ v_colors,
                    float* __restrict__ v_opacity) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= N) return;
    float4 g0 = make_float4(0.f, 0.f, 0.f, 0.f), g1 = g0, g2 = g0;
    if (__ldg(radii + i) > 0) {
        g0 = __ldg(grads + 3 * (size_t)i);
        g1 = __ldg(grads + 3 * (size_t)i + 1);
        g2 = __ldg(grads + 3 * (size_t)i + 2);
    }
    float a = __ldg(conics + 3 * i), b = __ldg(conics + 3 * i + 1), c = __ldg(conics + 3 * i + 2);
    // sigma = .5(a dx^2 + c dy^2) + b dx dy,  dx = x - px
    v_xys[i] = make_float2(a * g0.x + b * g0.y, b * g0.x + c * g0.y);
    v_conics[3 * i] = 0.5f * g0.z;
    v_conics[3 * i + 1] = g0.w;
    v_conics[3 * i + 2] = 0.5f * g1.x;
    v_opacity[i] = g1.y;
    v_colors[(size_t)CH * i] = g2.x;
    if (CH > 1) v_colors[(size_t)CH * i + 1] = g2.y;
    if (CH > 2) v_colors[(size_t)CH * i + 2] = g2.z;
    if (CH > 3) v_colors[(size_t)CH * i + 3] = g2.w;
}

// Data-parallel exchange: makes blend-backward's packed gradients self-contained
// before they leave the rank — rows of Gaussians culled in this view become exact zeros, the
// SH clamp mask [REF rasterize.py:39] is applied to the colour cotangents — and emits this view's
// d loss / d xy (the densification statistic is per view [REF model_gaussian.py:130-132]).
__global__ void __launch_bounds__(256)
dp_prepare_kernel(int N, const int32_t* __restrict__ radii, const uint8_t* __restrict__ clamp_mask,
                  const float4* __restrict__ recs, float4* __restrict__ grads, float2* __restrict__ v_xys) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= N) return;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!(__ldg(radii + i) > 0)) {
        grads[3 * (size_t)i] = zero;
        grads[3 * (size_t)i + 1] = zero;
        grads[3 * (size_t)i + 2] = zero;
        if (v_xys) v_xys[i] = make_float2(0.f, 0.f);
        return;
    }
    if (clamp_mask) {
        const unsigned m = clamp_mask[i];
        if ((m & 7u) != 7u) {
            float4 g2 = grads[3 * (size_t)i + 2];
            if (!(m & 1u)) g2.x = 0.f;
            if (!(m & 2u)) g2.y = 0.f;
            if (!(m & 4u)) g2.z = 0.f;
            grads[3 * (size_t)i + 2] = g2;
        }
    }
    if (v_xys) {
        const float4 g0 = grads[3 * (size_t)i];
        const float4 q1 = __ldg(recs + 3 * (size_t)i + 1);      // {.5 log2e a, log2e b, .5 log2e c, opacity}
        const float a = q1.x * (2.f / kLog2e), b = q1.y * (1.f / kLog2e), c = q1.z * (2.f / kLog2e);
        v_xys[i] = make_float2(a * g0.x + b * g0.y, b * g0.x + c * g0.y);
    }
}

#ifndef TS_HOST_EMU
// grouped backward (blend_group.cu)
int launch_blend_bwd_group(int CH, int gch, int H, int W, int tiles_x, int tiles_y,
                           const int32_t* tile_offsets, const int32_t* ids, const float* recs,
                           const float* background, const float* final_T, const int32_t* n_contrib,
                           const float* v_out_img, const float* v_out_ch3, int split_ch3,
                           const float* v_out_alpha, float* grads, const int32_t* order, cudaStream_t st);

// Which backward kernel ts_blend_bwd launches: 0 = first generation (blend_bwd_kernel, one warp
// per sub-block), 1 = grouped (blend_group.cu; default).  Initialised once from TS_BLEND_MODE
// ("warp" | "group"); ts_set_blend_mode() overrides it (tests, A/B benches).
static int g_blend_mode = -1;
static int blend_mode() {
    if (g_blend_mode < 0) {
        const char* e = getenv("TS_BLEND_MODE");
        if (e && !strcmp(e, "warp")) g_blend_mode = 0;
        else if (e && !strcmp(e, "group")) g_blend_mode = 1;
        else g_blend_mode = TS_DEFAULT_BLEND_MODE;
    }
    return g_blend_mode;
}
#endif  // !TS_HOST_EMU

}  // namespace ts

#ifndef TS_HOST_EMU
extern "C" {

int ts_set_blend_mode(int mode) {
    if (mode < -1 || mode > 1) return TS_ERR_INVALID;
    ts::g_blend_mode = mode;      // -1: back to TS_BLEND_MODE / the built-in default
    return TS_OK;
}
int ts_get_blend_mode(void) { return ts::blend_mode(); }

int ts_blend_fwd(int CH, int img_height, int img_width, int tiles_x, int tiles_y,
                 const int32_t* tile_offsets, const int32_t* ids_sorted, const float* recs,
                 const float* background, float* out_img, float* out_ch3, float* final_T,
                 int32_t* n_contrib, int clamp_max1, int capacity, const int32_t* tile_order,
                 ts_stream_t stream) {
    if (CH < 1 || CH > 4 || img_height <= 0 || img_width <= 0 || tiles_x <= 0 || tiles_y <= 0) return TS_ERR_INVALID;
    if (!tile_offsets || !background || !out_img || !final_T || !n_contrib) return TS_ERR_INVALID;
    if (recs && !ts::aligned16(recs)) return TS_ERR_ALIGN;
    dim3 grid(tiles_x * tiles_y);
    cudaStream_t st = (cudaStream_t)stream;
    const int cap = capacity > 0 ? capacity : INT32_MAX;
#define TS_LAUNCH_FWD(C) \
    ts::blend_fwd_kernel<C><<<grid, ts::kBlendThreads, 0, st>>>(img_height, img_width, tiles_x, tile_offsets, ids_sorted, (const float4*)recs, background, out_img, out_ch3, final_T, n_contrib, clamp_max1, cap, tile_order)
    switch (CH) {
        case 1: TS_LAUNCH_FWD(1); break;
        case 2: TS_LAUNCH_FWD(2); break;
        case 3: TS_LAUNCH_FWD(3); break;
        default: TS_LAUNCH_FWD(4); break;
    }
#undef TS_LAUNCH_FWD
    TS_CHECK_LAUNCH("ts_blend_fwd");
    return TS_OK;
}

int ts_blend_bwd(int N, int CH, int img_height, int img_width, int tiles_x, int tiles_y,
                 const int32_t* tile_offsets, const int32_t* ids_sorted, const float* recs,
                 const float* background, const float* final_T, const int32_t* n_contrib,
                 const float* v_out_img, const float* v_out_ch3, int split_ch3_flags,
                 const float* v_out_alpha, float* grads, const int32_t* tile_order, ts_stream_t stream) {
    const int split_ch3 = split_ch3_flags & 1;
    const bool prezeroed = (split_ch3_flags & TS_BLEND_GRADS_ZEROED) != 0;
    if (N < 0 || CH < 1 || CH > 4 || img_height <= 0 || img_width <= 0 || tiles_x <= 0 || tiles_y <= 0) return TS_ERR_INVALID;
    if (N == 0) return TS_OK;
    if (!tile_offsets || !background || !final_T || !n_contrib || !grads) return TS_ERR_INVALID;
    if (!v_out_img && !(CH == 4 && split_ch3)) return TS_ERR_INVALID;
    if (!ts::aligned16(grads) || (recs && !ts::aligned16(recs))) return TS_ERR_ALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    if (!prezeroed)
        TS_CHECK_CUDA(cudaMemsetAsync(grads, 0, sizeof(float) * ts::kGradFloats * (size_t)N, st), "ts_blend_bwd/memset");
    dim3 grid(tiles_x * tiles_y);
    // channels that carry a cotangent: the fused RGB+depth pass without a depth loss skips ch 3
    const int gch = (CH == 4 && split_ch3 && !v_out_ch3) ? 3 : CH;
    if (ts::blend_mode() == 1) {
        ts::launch_blend_bwd_group(CH, gch, img_height, img_width, tiles_x, tiles_y, tile_offsets, ids_sorted,
                                   recs, background, final_T, n_contrib, v_out_img, v_out_ch3, split_ch3,
                                   v_out_alpha, grads, tile_order, st);
        TS_CHECK_LAUNCH("ts_blend_bwd/group");
        return TS_OK;
    }
#define TS_LAUNCH_BWD(C, G)                                                                              \
    do {                                                                                                 \
        /* per device and context, cheap: set on every call (a process may drive several GPUs) */       \
        TS_CHECK_CUDA(cudaFuncSetAttribute(ts::blend_bwd_kernel<C, G>,                                   \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize,                  \
                                           (int)ts::kBwdSmemBytes), "ts_blend_bwd/attr");                \
        ts::blend_bwd_kernel<C, G><<<grid, ts::kBlendThreads, ts::kBwdSmemBytes, st>>>(                  \
            img_height, img_width, tiles_x, tile_offsets, ids_sorted, (const float4*)recs, background,   \
            final_T, n_contrib, v_out_img, v_out_ch3, split_ch3, v_out_alpha, (float4*)grads, tile_order); \
    } while (0)
    switch (CH) {
        case 1: TS_LAUNCH_BWD(1, 1); break;
        case 2: TS_LAUNCH_BWD(2, 2); break;
        case 3: TS_LAUNCH_BWD(3, 3); break;
        default:
            if (gch == 3) TS_LAUNCH_BWD(4, 3); else TS_LAUNCH_BWD(4, 4);
            break;
    }
#undef TS_LAUNCH_BWD
    TS_CHECK_LAUNCH("ts_blend_bwd");
    return TS_OK;
}

int ts_blend_unpack_grads(int N, int CH, const int32_t* radii, const float* conics,
                          const float* grads, float* v_xys, float* v_conics, float* v_colors,
                          float* v_opacity, ts_stream_t stream) {
    if (N < 0 || CH < 1 || CH > 4) return TS_ERR_INVALID;
    if (N == 0) return TS_OK;
    if (!radii || !conics || !grads || !v_xys || !v_conics || !v_colors || !v_opacity) return TS_ERR_INVALID;
    if (!ts::aligned16(grads) || (reinterpret_cast<uintptr_t>(v_xys) & 7u)) return TS_ERR_ALIGN;
    int grid = (N + 255) / 256;
    cudaStream_t st = (cudaStream_t)stream;
#define TS_LAUNCH_UNPACK(C) \
    ts::unpack_grads_kernel<C><<<grid, 256, 0, st>>>(N, radii, conics, (const float4*)grads, (float2*)v_xys, v_conics, v_colors, v_opacity)
    switch (CH) {
        case 1: TS_LAUNCH_UNPACK(1); break;
        case 2: TS_LAUNCH_UNPACK(2); break;
        case 3: TS_LAUNCH_UNPACK(3); break;
        default: TS_LAUNCH_UNPACK(4); break;
    }
#undef TS_LAUNCH_UNPACK
    TS_CHECK_LAUNCH("ts_blend_unpack_grads");
    return TS_OK;
}

int ts_dp_prepare(int N, const int32_t* radii, const uint8_t* clamp_mask, const float* recs, float* grads,
                  float* v_xys, ts_stream_t stream) {
    if (N < 0) return TS_ERR_INVALID;
    if (N == 0) return TS_OK;
    if (!radii || !recs || !grads) return TS_ERR_INVALID;
    if (!ts::aligned16(recs) || !ts::aligned16(grads) || (v_xys && (reinterpret_cast<uintptr_t>(v_xys) & 7u)))
        return TS_ERR_ALIGN;
    ts::dp_prepare_kernel<<<(N + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
        N, radii, clamp_mask, (const float4*)recs, (float4*)grads, (float2*)v_xys);
    TS_CHECK_LAUNCH("ts_dp_prepare");
    return TS_OK;
}

}  // extern "C"
#endif  // !TS_HOST_EMU
