// K5, second generation — "grouped" blend-backward (default; the first generation in blend.cu stays
// selectable with ts_set_blend_mode(0) / TS_BLEND_MODE=warp for A/B runs).
// Behind gsplat.rasterize_gaussians  [REF tinysplat/splatting/rasterize.py:44,50,83-86].
//
// Mapping.  One CTA of 64 threads per 16x16 tile [REF rasterize.py:19-20].  The 64 threads form
// eight 8-lane GROUPS; group s owns the 8x4-pixel sub-block s of the tile, lane = column, and
// every lane keeps the compositing state of its FOUR rows in registers.  The four groups of a
// warp walk four DIFFERENT candidate lists in lock-step (per-lane shared-memory addresses), so
// one warp-instruction evaluates one pixel row of four (sub-block, Gaussian) pairs.
//
// Why (tests/analysis/cull_stats.py on synthetic_1M_1080p; measured history in profiles/README.md):
//   * culling is exact per pixel row (footprint_rowmask: the interval of each row the ellipse
//     {alpha >= 1/255} covers), not a bounding box per sub-block: 6.15 M -> 5.31 M
//     (sub-block, Gaussian) pairs;
//   * a lane first sums its four rows in registers (a lane is a pixel column, so dx is shared and
//     three moments per row suffice), then ONE 8-lane transpose-reduce finishes a pair — the
//     first generation paid a 32-lane reduction through shared memory for every pair;
//   * the colour accumulated behind a Gaussian only enters through its dot product with the
//     pixel's fixed cotangent: one scalar per pixel replaces the per-channel buffers;
//   * group totals go straight to global memory as red.global.add.v4/.v2.f32 from two lanes of
//     the group: fire-and-forget, no shared accumulators, no flush phase (shared memory has no
//     native fp32 add: the accumulator variant spent 27 % of its stall samples in CAS loops);
//   * 64 registers, 14 KB shared memory -> 15 CTAs (30 warps) per SM.
// Still not HBM-bound: fp32 FMA / MUFU issue.
#include "ts_blend_common.cuh"
#include "ts_f32x2.cuh"

namespace ts {

constexpr int kGThreads = 64;                  // 2 warps = 8 groups of 8 lanes
#ifndef TS_GBATCH
#define TS_GBATCH 128
#endif
constexpr int kGBatch = TS_GBATCH;             // candidates staged per batch (64, 128 or 256)
constexpr int kGPer = kGBatch / kGThreads;     // records gathered per thread per batch
constexpr int kGWords = kGBatch / 32;          // candidate-mask words per group per batch

struct GroupMap {
    int lane, warp, grp, l8, shift;
    int j, i0;          // pixel column, first of the four rows
    float px, py0;
    unsigned inside;    // bit r: pixel (i0 + r, j) is inside the image
    float X0, Y0;       // pixel centre of the tile's first pixel
};

__device__ __forceinline__ GroupMap group_map(int H, int W, int bx, int by) {
    GroupMap m;
    const int tid = threadIdx.x;
    m.lane = tid & 31;
    m.warp = tid >> 5;
    m.grp = tid >> 3;                    // sub-block id: wx = grp & 1, wy = grp >> 1
    m.l8 = tid & 7;
    const int wx = m.grp & 1, wy = m.grp >> 1;
    m.shift = 8 * wy + wx;               // rowmask bit of (row 4wy + r, half wx) = shift + 2r
    m.j = bx * kBlock + wx * 8 + m.l8;
    m.i0 = by * kBlock + wy * 4;
    m.px = (float)m.j + kPixCenter;
    m.py0 = (float)m.i0 + kPixCenter;
    m.inside = 0u;
#pragma unroll
    for (int r = 0; r < 4; ++r)
        if (m.j < W && m.i0 + r < H) m.inside |= 1u << r;
    m.X0 = (float)(bx * kBlock) + kPixCenter;
    m.Y0 = (float)(by * kBlock) + kPixCenter;
    return m;
}

// Stage-time masks of one batch: s_rmask[t] = exact row mask of staged record t,
// s_cmask[s * kGWords + k] = which of records 32k..32k+31 can reach sub-block s.
// `valid[jj]`: this thread's record jj of the batch exists.  Must be reached by all 64 threads.
__device__ __forceinline__ void build_masks(const GroupMap& gm, const float4* rec, const bool (&valid)[kGPer],
                                            unsigned* s_rmask, unsigned* s_cmask) {
    const unsigned full = 0xffffffffu;
#pragma unroll
    for (int jj = 0; jj < kGPer; ++jj) {
        const int t = jj * kGThreads + threadIdx.x;
        unsigned rm = 0u;
        if (valid[jj]) rm = footprint_rowmask(rec[t * 3], rec[t * 3 + 1], gm.X0, gm.Y0);
        s_rmask[t] = rm;
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            const unsigned sel = 0x55u << (8 * (s >> 1) + (s & 1));
            const unsigned w = __ballot_sync(full, (rm & sel) != 0u);
            if (gm.lane == 0) s_cmask[s * kGWords + jj * 2 + gm.warp] = w;
        }
    }
}

// Sums val[0..NV) over the 8 lanes of each group.  Level xor 4 transposes (8 -> 4 values per
// lane), levels xor 2 and xor 1 are plain butterflies, so lanes with l8 < 4 end with the group
// totals of values 0..3 and lanes with l8 >= 4 with those of values 4..7 — whole float4s of the
// packed gradient record.  Values 8, 9 are reduced plainly and returned on every lane in `extra`.
template <int NV>
__device__ __forceinline__ float4 group_reduce_quad(const float (&val)[NV], int l8, float2& extra) {
    const unsigned full = 0xffffffffu;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = (i < NV) ? val[i < NV ? i : 0] : 0.f;
    const bool h4 = (l8 & 4) != 0;
    float keep[4], recv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float send = h4 ? v[i] : v[i + 4];
        keep[i] = h4 ? v[i + 4] : v[i];
        recv[i] = __shfl_xor_sync(full, send, 4);
    }
    // the four running sums live in two packed pairs: one FADD2 adds both halves
    f32x2 w01 = add2(pack2(keep[0], keep[1]), pack2(recv[0], recv[1]));
    f32x2 w23 = add2(pack2(keep[2], keep[3]), pack2(recv[2], recv[3]));
#pragma unroll
    for (int d = 2; d > 0; d >>= 1) {
        const float a = __shfl_xor_sync(full, lo2(w01), d), b = __shfl_xor_sync(full, hi2(w01), d);
        const float c = __shfl_xor_sync(full, lo2(w23), d), e = __shfl_xor_sync(full, hi2(w23), d);
        w01 = add2(w01, pack2(a, b));
        w23 = add2(w23, pack2(c, e));
    }
    float e[2] = {0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        if (8 + k < NV) {
            float x = val[8 + k < NV ? 8 + k : 0];
            x += __shfl_xor_sync(full, x, 4);
            x += __shfl_xor_sync(full, x, 2);
            x += __shfl_xor_sync(full, x, 1);
            e[k] = x;
        }
    }
    extra = make_float2(e[0], e[1]);
    return make_float4(lo2(w01), hi2(w01), lo2(w23), hi2(w23));
}

// GCH = number of colour channels that carry a cotangent (GCH <= CH; the fused RGB+depth pass
// with no depth loss has CH = 4, GCH = 3 and skips all channel-3 gradient arithmetic).
// 16 CTAs per SM = a 64-register budget (measured: the packed-pair loop compiles to 72 registers without
// the bound and is 6 % slower at the lower occupancy)
template <int CH, int GCH>
__global__ void __launch_bounds__(kGThreads, 16)
blend_bwd_group_kernel(int H, int W, int tbx, const int32_t* __restrict__ tile_offsets,
                       const int32_t* __restrict__ ids, const float4* __restrict__ recs,
                       const float* __restrict__ background, const float* __restrict__ final_T,
                       const int32_t* __restrict__ n_contrib, const float* __restrict__ v_out_img,
                       const float* __restrict__ v_out_ch3, int split_ch3,
                       const float* __restrict__ v_out_alpha, float4* __restrict__ grads,
                       const int32_t* __restrict__ order) {
    constexpr int NV = 6 + GCH;    // values reduced per (sub-block, Gaussian) pair
    __shared__ __align__(16) float4 s_rec[2][kGBatch * 3];
    __shared__ int s_gid[2 * kGBatch];                             // gaussian ids of the staged records
    __shared__ unsigned s_rmask[kGBatch];
    __shared__ unsigned s_cmask[8 * kGWords];
    __shared__ int s_nmax;
    const unsigned full = 0xffffffffu;
    const TileId tl = tile_id(order, tbx);
    const GroupMap gm = group_map(H, W, tl.bx, tl.by);
    const int tid = threadIdx.x;
    const int tile = tl.tile;
    const int start = __ldg(tile_offsets + tile);
    const unsigned gbits = 0xffu << (gm.lane & 24);      // the lanes of my group

    // per-pixel state of this lane's four rows
    // W = T_final (v_alpha_out - bg.v) - sum_c buffer_c v_c: the colour accumulated BEHIND the
    // current Gaussian only ever enters through its dot product with the pixel's (fixed)
    // cotangent, so one scalar per pixel replaces the per-channel buffers.
    float T[4], Wacc[4], v_out[4][GCH];
    int nc[4];
    int ncmax = 0;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        float T_final = 1.f, v_oa = 0.f;
        nc[r] = 0;
#pragma unroll
        for (int c = 0; c < GCH; ++c) v_out[r][c] = 0.f;
        if ((gm.inside >> r) & 1u) {
            const size_t pix = (size_t)(gm.i0 + r) * W + gm.j;
            T_final = __ldg(final_T + pix);
            int n = __ldg(n_contrib + pix);
            const unsigned cm = (unsigned)n >> kClampShift;    // channels clamped by forward
            nc[r] = n & kCountMask;
            if (CH == 4 && split_ch3) {
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    v_out[r][c] = (v_out_img && !((cm >> c) & 1u)) ? __ldg(v_out_img + pix * 3 + c) : 0.f;
                if (GCH == 4) v_out[r][GCH - 1] = v_out_ch3 ? __ldg(v_out_ch3 + pix) : 0.f;
            } else {
#pragma unroll
                for (int c = 0; c < GCH; ++c) v_out[r][c] = __ldg(v_out_img + pix * CH + c);
            }
            if (v_out_alpha) v_oa = __ldg(v_out_alpha + pix);
        }
        float bgdot = 0.f;
#pragma unroll
        for (int c = 0; c < GCH; ++c) bgdot = fmaf(__ldg(background + c), v_out[r][c], bgdot);
        Wacc[r] = T_final * (v_oa - bgdot);
        T[r] = T_final;
        ncmax = max(ncmax, nc[r]);
    }

    // rows (0,1) and (2,3) of the column as packed pairs: every per-row fp32 operation of the loop
    // below is issued once per PAIR (FFMA2 / FMUL2 / FADD2)
    f32x2 T2[2], W2[2], vo2[2][GCH], py2[2];
#pragma unroll
    for (int P = 0; P < 2; ++P) {
        T2[P] = pack2(T[2 * P], T[2 * P + 1]);
        W2[P] = pack2(Wacc[2 * P], Wacc[2 * P + 1]);
        py2[P] = pack2(gm.py0 + (float)(2 * P), gm.py0 + (float)(2 * P + 1));
#pragma unroll
        for (int c = 0; c < GCH; ++c) vo2[P][c] = pack2(v_out[2 * P][c], v_out[2 * P + 1][c]);
    }

    if (tid == 0) s_nmax = 0;
    __syncthreads();
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) ncmax = max(ncmax, __shfl_xor_sync(full, ncmax, d));
    if (gm.lane == 0) atomicMax(&s_nmax, ncmax);
    __syncthreads();
    const int nmax = s_nmax;   // entries [0, nmax) of the tile list contributed somewhere
    const int nb = (nmax + kGBatch - 1) / kGBatch;

    // batch b, slot t  <->  list position  p = nmax-1 - (b*kGBatch + t)   (back to front)
    auto prefetch = [&](int b) {
#pragma unroll
        for (int jj = 0; jj < kGPer; ++jj) {
            const int t = jj * kGThreads + tid;
            const int p = nmax - 1 - (b * kGBatch + t);
            if (p >= 0) {
                const int g = __ldg(ids + start + p);
                s_gid[(b & 1) * kGBatch + t] = g;
                const float4* src = recs + 3 * (size_t)g;
                float4* dst = &s_rec[b & 1][t * 3];
                cp_async16(dst, src);
                cp_async16(dst + 1, src + 1);
                cp_async16(dst + 2, src + 2);
            }
        }
    };
    if (nb > 0) prefetch(0);
    cp_async_commit();

    for (int b = 0; b < nb; ++b) {
        const float4* rec = s_rec[b & 1];
        bool valid[kGPer];
#pragma unroll
        for (int jj = 0; jj < kGPer; ++jj) valid[jj] = nmax - 1 - (b * kGBatch + jj * kGThreads + tid) >= 0;
        if (b + 1 < nb) prefetch(b + 1);
        cp_async_commit();
        cp_async_wait<1>();
        build_masks(gm, rec, valid, s_rmask, s_cmask);
        __syncthreads();
        const int pbase = nmax - 1 - b * kGBatch;

        int k = 0;
        unsigned m = s_cmask[gm.grp * kGWords];
        for (;;) {
            while (m == 0u && k < kGWords - 1) m = s_cmask[gm.grp * kGWords + (++k)];
            const bool act = (m != 0u);
            if (!__any_sync(full, act)) break;
            int c = 0, gid = 0;
            float val[NV];
#pragma unroll
            for (int v = 0; v < NV; ++v) val[v] = 0.f;
            bool any = false;
            if (act) {
                c = k * 32 + __ffs(m) - 1;
                m &= m - 1;
                const int p = pbase - c;
                gid = s_gid[(b & 1) * kGBatch + c];     // loaded here: its latency hides behind the rows
                const unsigned rm = s_rmask[c] >> gm.shift;
                const float4 q0 = rec[c * 3];
                const float4 q1 = rec[c * 3 + 1];
                const float4 q2 = rec[c * 3 + 2];
                const float col[4] = {q2.x, q2.y, q2.z, q2.w};
                // branch-free over the four rows, two at a time.  A lane is one pixel COLUMN, so dx is
                // common to its rows and only s0 = sum v_sigma, s1 = sum v_sigma dy, s2 = sum v_sigma dy^2
                // are accumulated per row; the five geometric sums follow from them below, and
                // v_opacity = sum vis * v_alpha = -s0 / opacity  (v_sigma = -opacity vis v_alpha).
                // The accumulators hold -s0, -s1, -s2 (the sign is applied once, after the rows).
                const float dxc = __fsub_rn(q0.x, gm.px);
                const float axd = __fmul_rn(q1.x, dxc);
                f32x2 n0, n1, n2, vc[GCH];
#pragma unroll
                for (int P = 0; P < 2; ++P) {
                    f32x2 dy;
                    const f32x2 pw = eval_power2(dxc, axd, q0.y, q1.y, q1.z, py2[P], dy);
                    const float pwa = lo2(pw), pwb = hi2(pw);
                    const f32x2 araw = mul2(bcast2(q1.w), pack2(ex2_approx(-fmaxf(pwa, 0.f)), ex2_approx(-fmaxf(pwb, 0.f))));
                    const float ara = lo2(araw), arb = hi2(araw);
                    const float ama = fminf(kAlphaMax, ara), amb = fminf(kAlphaMax, arb);
                    const bool oka = (rm & (1u << (4 * P))) != 0u && p < nc[2 * P] && pwa >= 0.f && ama >= kAlphaMin;
                    const bool okb = (rm & (4u << (4 * P))) != 0u && p < nc[2 * P + 1] && pwb >= 0.f && amb >= kAlphaMin;
                    any = any || oka || okb;
                    const f32x2 alpha = pack2(oka ? ama : 0.f, okb ? amb : 0.f);
                    const f32x2 oma = sub2(bcast2(1.f), alpha);
                    // rcp_approx(1) == 1 exactly (checked on the device by the tests): a row that does
                    // not contribute keeps its T
                    const f32x2 ra = pack2(rcp_approx(lo2(oma)), rcp_approx(hi2(oma)));
                    T2[P] = mul2(T2[P], ra);
                    const f32x2 fac = mul2(alpha, T2[P]);                  // 0 when !ok
                    f32x2 cv = mul2(bcast2(col[0]), vo2[P][0]);
#pragma unroll
                    for (int ch = 1; ch < GCH; ++ch) cv = fma2(bcast2(col[ch]), vo2[P][ch], cv);
#pragma unroll
                    for (int ch = 0; ch < GCH; ++ch) vc[ch] = (P == 0) ? mul2(fac, vo2[P][ch]) : fma2(fac, vo2[P][ch], vc[ch]);
                    const f32x2 v_alpha = fma2(cv, T2[P], mul2(W2[P], ra));
                    W2[P] = fma2(neg2(cv), fac, W2[P]);
                    // !ok: no contribution;  clamped alpha: d alpha / d araw = 0
                    const f32x2 ng = mul2(pack2((oka && !(ara > kAlphaMax)) ? ara : 0.f,
                                                (okb && !(arb > kAlphaMax)) ? arb : 0.f), v_alpha);   // -v_sigma
                    const f32x2 ngdy = mul2(ng, dy);
                    n0 = (P == 0) ? ng : add2(n0, ng);
                    n1 = (P == 0) ? ngdy : add2(n1, ngdy);
                    n2 = (P == 0) ? mul2(ngdy, dy) : fma2(ngdy, dy, n2);
                }
                const float s0 = -hsum2(n0), s1 = -hsum2(n1), s2 = -hsum2(n2);
#pragma unroll
                for (int ch = 0; ch < GCH; ++ch) val[6 + ch] = hsum2(vc[ch]);
                val[0] = s0 * dxc;
                val[1] = s1;
                val[2] = val[0] * dxc;
                val[3] = s1 * dxc;
                val[4] = s2;
                val[5] = any ? -s0 * rcp_approx(q1.w) : 0.f;   // any => opacity >= 1/255
            }
            const unsigned anyb = __ballot_sync(full, any);
            if (anyb == 0u) continue;
            float2 extra;
            const float4 q = group_reduce_quad<NV>(val, gm.l8, extra);
            // lane 0 of a group holds packed floats 0..3; lane 4 holds values 4..7 = floats 4, 5
            // (S_yy, v_opacity) and the first two colours, which with the extras form g2
            if (act && (anyb & gbits) != 0u && (gm.l8 & 3) == 0) {
                float4* dst = grads + 3 * (size_t)gid;
                if (gm.l8 == 0) {
                    atomicAdd(dst, q);
                } else {
                    atomicAdd(reinterpret_cast<float2*>(dst + 1), make_float2(q.x, q.y));
                    atomicAdd(dst + 2, make_float4(q.z, q.w, extra.x, extra.y));
                }
            }
        }
        __syncthreads();  // buffers free before the next batch touches them
    }
    cp_async_wait<0>();
}

#ifndef TS_HOST_EMU
// ---- launchers (called from the C-ABI entry points in blend.cu) ------------------------------
int launch_blend_bwd_group(int CH, int gch, int H, int W, int tiles_x, int tiles_y,
                           const int32_t* tile_offsets, const int32_t* ids, const float* recs,
                           const float* background, const float* final_T, const int32_t* n_contrib,
                           const float* v_out_img, const float* v_out_ch3, int split_ch3,
                           const float* v_out_alpha, float* grads, const int32_t* order, cudaStream_t st) {
    dim3 grid(tiles_x * tiles_y);
#define TS_LAUNCH_BWD(C, G)                                                                        \
    blend_bwd_group_kernel<C, G><<<grid, kGThreads, 0, st>>>(                                      \
        H, W, tiles_x, tile_offsets, ids, (const float4*)recs, background, final_T, n_contrib,     \
        v_out_img, v_out_ch3, split_ch3, v_out_alpha, (float4*)grads, order)
    switch (CH) {
        case 1: TS_LAUNCH_BWD(1, 1); break;
        case 2: TS_LAUNCH_BWD(2, 2); break;
        case 3: TS_LAUNCH_BWD(3, 3); break;
        default:
            if (gch == 3) TS_LAUNCH_BWD(4, 3); else TS_LAUNCH_BWD(4, 4);
            break;
    }
#undef TS_LAUNCH_BWD
    return 0;
}

// test hook: the device's approximate-math results the blend kernels build on
__global__ void debug_approx_kernel(int n, const float* __restrict__ x, float* __restrict__ rcp, float* __restrict__ ex2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { rcp[i] = rcp_approx(x[i]); ex2[i] = ex2_approx(x[i]); }
}
#endif  // !TS_HOST_EMU

}  // namespace ts

#ifndef TS_HOST_EMU
extern "C" {

int ts_debug_approx(int n, const float* x, float* rcp_out, float* ex2_out, ts_stream_t stream) {
    if (n <= 0 || !x || !rcp_out || !ex2_out) return TS_ERR_INVALID;
    ts::debug_approx_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(n, x, rcp_out, ex2_out);
    TS_CHECK_LAUNCH("ts_debug_approx");
    return TS_OK;
}

// Host-side evaluation of the exact row mask of one packed record against tile (tile_x, tile_y):
// the same function the kernels run (test hook: tests/test_capi.py brute-forces it on CPU).
uint32_t ts_debug_rowmask(const float* q0, const float* q1, int tile_x, int tile_y) {
    const float4 a = make_float4(q0[0], q0[1], q0[2], q0[3]);
    const float4 b = make_float4(q1[0], q1[1], q1[2], q1[3]);
    return ts::footprint_rowmask(a, b, (float)(tile_x * ts::kBlock) + ts::kPixCenter,
                                 (float)(tile_y * ts::kBlock) + ts::kPixCenter);
}


}  // extern "C"
#endif  // !TS_HOST_EMU
