// K2 / K7 — spherical harmonics -> colour, forward and backward (streaming, HBM-bound).
// Replaces gsplat.sh.spherical_harmonics  [REF tinysplat/splatting/rasterize.py:38,75-81].
// One thread per Gaussian; the [N,K,3] coefficient rows are staged through shared memory with
// coalesced (128-bit where alignment allows) accesses, padded to an odd stride so the
// thread-per-row reads are bank-conflict free.  Only the first (degree+1)^2 bases are read.
#include "ts_common.cuh"
#include "ts_sh_basis.cuh"

namespace ts {

constexpr int kShThreads = 128;

// Copy the first `need` floats of each `row`-float row of items [item0, item0+n_valid) into
// shared rows of stride `sstride` at column `scol`.  128-bit global loads when rows keep 16 B
// alignment and the needed prefix is a multiple of 4 floats.
__device__ __forceinline__ void rows_to_smem(const float* __restrict__ g, int row, int need,
                                             float* s, int sstride, int scol, int item0,
                                             int n_valid) {
    const float* base = g + (size_t)item0 * row;
    if (((row | need) & 3) == 0 && aligned_dev16(base)) {
        int nv = need >> 2, total = n_valid * nv;
        for (int idx = threadIdx.x; idx < total; idx += kShThreads) {
            int it = idx / nv, off = (idx - it * nv) << 2;
            float4 v = __ldg(reinterpret_cast<const float4*>(base + (size_t)it * row + off));
            float* d = s + it * sstride + scol + off;
            d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
        }
    } else {
        int total = n_valid * need;
        for (int idx = threadIdx.x; idx < total; idx += kShThreads) {
            int it = idx / need, off = idx - it * need;
            s[it * sstride + scol + off] = __ldg(base + (size_t)it * row + off);
        }
    }
}

// Write full `row`-float rows from shared (stride sstride, column scol) to global.
__device__ __forceinline__ void smem_to_rows(float* __restrict__ g, int row, const float* s,
                                             int sstride, int scol, int item0, int n_valid) {
    float* base = g + (size_t)item0 * row;
    if ((row & 3) == 0 && aligned_dev16(base)) {
        int nv = row >> 2, total = n_valid * nv;
        for (int idx = threadIdx.x; idx < total; idx += kShThreads) {
            int it = idx / nv, off = (idx - it * nv) << 2;
            const float* d = s + it * sstride + scol + off;
            *reinterpret_cast<float4*>(base + (size_t)it * row + off) = make_float4(d[0], d[1], d[2], d[3]);
        }
    } else {
        int total = n_valid * row;
        for (int idx = threadIdx.x; idx < total; idx += kShThreads) {
            int it = idx / row, off = idx - it * row;
            base[(size_t)it * row + off] = s[it * sstride + scol + off];
        }
    }
}

// Unit view direction of item `tid`: either given, or derived from the mean and the view
// matrix' translation column exactly as the reference adapter does [REF rasterize.py:77-79].
__device__ __forceinline__ void load_dir(const float* s_dir, int tid, int flags,
                                         const float* __restrict__ viewmat, float& dx, float& dy,
                                         float& dz) {
    dx = s_dir[3 * tid]; dy = s_dir[3 * tid + 1]; dz = s_dir[3 * tid + 2];
    if (flags & TS_SH_DIRS_FROM_MEANS) {
        dx -= __ldg(viewmat + 3); dy -= __ldg(viewmat + 7); dz -= __ldg(viewmat + 11);
    }
}

template <int DEG>
__global__ void __launch_bounds__(kShThreads)
sh_fwd_kernel(int N, int K, const float* __restrict__ dirs, const float* __restrict__ viewmat,
              const float* __restrict__ coeffs, const float* __restrict__ coeffs_rest,
              float* __restrict__ colors, int out_stride, const float* __restrict__ ch3,
              uint8_t* __restrict__ clamp_mask, int flags, int sstride) {
    TS_DYN_SMEM(float, s_sh, 16);
    constexpr int NB = (DEG + 1) * (DEG + 1);
    float* s_dir = s_sh;                        // [TH*3], reused for the colour output
    float* s_co = s_sh + kShThreads * 3 + 4;    // [TH][sstride]
    const int item0 = blockIdx.x * kShThreads;
    const int n_valid = min(kShThreads, N - item0);
    const int tid = threadIdx.x;
    block_load<3, kShThreads>(dirs, s_dir, item0, N);
    if (coeffs_rest == nullptr) {
        rows_to_smem(coeffs, K * 3, NB * 3, s_co, sstride, 0, item0, n_valid);
    } else {
        rows_to_smem(coeffs, 3, 3, s_co, sstride, 0, item0, n_valid);
        if (NB > 1) rows_to_smem(coeffs_rest, (K - 1) * 3, (NB - 1) * 3, s_co, sstride, 3, item0, n_valid);
    }
    __syncthreads();
    float r = 0.f, g = 0.f, bl = 0.f;
    if (tid < n_valid) {
        float b[NB];
        float dx, dy, dz;
        load_dir(s_dir, tid, flags, viewmat, dx, dy, dz);
        sh_basis<DEG>(dx, dy, dz, b);
        const float* c = s_co + tid * sstride;
#pragma unroll
        for (int k = 0; k < NB; ++k) {
            r = fmaf(b[k], c[3 * k], r);
            g = fmaf(b[k], c[3 * k + 1], g);
            bl = fmaf(b[k], c[3 * k + 2], bl);
        }
        if (flags & TS_SH_OFFSET_CLAMP) {   // clamp(rgb + 0.5, min=0) [REF rasterize.py:39]
            r += 0.5f; g += 0.5f; bl += 0.5f;
            unsigned m = (r >= 0.f ? 1u : 0u) | (g >= 0.f ? 2u : 0u) | (bl >= 0.f ? 4u : 0u);
            r = fmaxf(r, 0.f); g = fmaxf(g, 0.f); bl = fmaxf(bl, 0.f);
            if (clamp_mask) clamp_mask[item0 + tid] = (uint8_t)m;
        }
    }
    if (out_stride == 3) {
        __syncthreads();
        s_dir[3 * tid] = r; s_dir[3 * tid + 1] = g; s_dir[3 * tid + 2] = bl;
        __syncthreads();
        block_store<3, kShThreads>(colors, s_dir, item0, N);
    } else if (tid < n_valid) {
        // strided output (e.g. straight into the third float4 of the packed raster record);
        // channel 3 carries `ch3` (the depth, in the fused RGB+depth pass)
        float* o = colors + (size_t)out_stride * (item0 + tid);
        float d = ch3 ? __ldg(ch3 + item0 + tid) : 0.f;
        if ((out_stride & 3) == 0 && aligned_dev16(colors)) {
            *reinterpret_cast<float4*>(o) = make_float4(r, g, bl, d);
        } else {
            o[0] = r; o[1] = g; o[2] = bl;
            if (ch3) o[3] = d;
        }
    }
}

template <int DEG>
__global__ void __launch_bounds__(kShThreads)
sh_bwd_kernel(int N, int K, const float* __restrict__ dirs, const float* __restrict__ viewmat,
              const float* __restrict__ v_colors, int v_stride,
              const uint8_t* __restrict__ clamp_mask, float* __restrict__ v_coeffs,
              float* __restrict__ v_coeffs_rest, int flags, int sstride) {
    TS_DYN_SMEM(float, s_sh, 16);
    constexpr int NB = (DEG + 1) * (DEG + 1);
    float* s_dir = s_sh;                          // [TH*3]
    float* s_vc = s_sh + kShThreads * 3 + 4;      // [TH*3]
    float* s_co = s_sh + 2 * (kShThreads * 3 + 4);  // [TH][sstride]
    const int item0 = blockIdx.x * kShThreads;
    const int n_valid = min(kShThreads, N - item0);
    const int tid = threadIdx.x;
    block_load<3, kShThreads>(dirs, s_dir, item0, N);
    if (v_stride == 3) block_load<3, kShThreads>(v_colors, s_vc, item0, N);
    __syncthreads();
    if (tid < n_valid) {
        float b[NB];
        float dx, dy, dz;
        load_dir(s_dir, tid, flags, viewmat, dx, dy, dz);
        sh_basis<DEG>(dx, dy, dz, b);
        float vr, vg, vb;
        if (v_stride == 3) {
            vr = s_vc[3 * tid]; vg = s_vc[3 * tid + 1]; vb = s_vc[3 * tid + 2];
        } else {
            const float* v = v_colors + (size_t)v_stride * (item0 + tid);
            if ((v_stride & 3) == 0 && aligned_dev16(v_colors)) {
                float4 t = __ldg(reinterpret_cast<const float4*>(v));
                vr = t.x; vg = t.y; vb = t.z;
            } else {
                vr = __ldg(v); vg = __ldg(v + 1); vb = __ldg(v + 2);
            }
        }
        if (clamp_mask) {
            unsigned m = clamp_mask[item0 + tid];
            if (!(m & 1u)) vr = 0.f;
            if (!(m & 2u)) vg = 0.f;
            if (!(m & 4u)) vb = 0.f;
        }
        float* c = s_co + tid * sstride;
#pragma unroll
        for (int k = 0; k < NB; ++k) {
            c[3 * k] = b[k] * vr;
            c[3 * k + 1] = b[k] * vg;
            c[3 * k + 2] = b[k] * vb;
        }
        for (int k = NB * 3; k < K * 3; ++k) c[k] = 0.f;
    }
    __syncthreads();
    if (v_coeffs_rest == nullptr) {
        smem_to_rows(v_coeffs, K * 3, s_co, sstride, 0, item0, n_valid);
    } else {
        smem_to_rows(v_coeffs, 3, s_co, sstride, 0, item0, n_valid);
        if (K > 1) smem_to_rows(v_coeffs_rest, (K - 1) * 3, s_co, sstride, 3, item0, n_valid);
    }
}

// ---- TMA variants for the fused pipeline ------------------------------------------------------
// Coefficients arrive as (dc[N,3], rest[N,K-1,3]) and the block's rows are one contiguous chunk
// of global memory, so a single elected thread moves them with cp.async.bulk (UBLKCP) straight
// into a DENSE shared layout: row stride (K-1)*3 = 45 floats at degree 3 is odd, i.e. already
// bank-conflict free for the thread-per-row reads — no padding, no per-element index math.
// The last (partial) block falls back to a plain coalesced copy of the same dense chunk.
template <int THREADS>
__device__ __forceinline__ void dense_copy_in(float* s, const float* __restrict__ g, int nfl) {
    for (int i = threadIdx.x; i < nfl; i += THREADS) s[i] = __ldg(g + i);
}

template <int DEG>
__global__ void __launch_bounds__(kShThreads)
sh_fwd_bulk_kernel(int N, int K, const float* __restrict__ means, const float* __restrict__ viewmat,
                   const float* __restrict__ dc, const float* __restrict__ rest,
                   float* __restrict__ colors, int out_stride, const float* __restrict__ ch3,
                   uint8_t* __restrict__ clamp_mask, int flags) {
    TS_DYN_SMEM(float, s_sh, 128);
    constexpr int NB = (DEG + 1) * (DEG + 1);
    const int R = (K - 1) * 3;                       // floats per `rest` row
    float* s_rest = s_sh;                            // [TH][R] dense
    float* s_dc = s_rest + kShThreads * R;           // [TH*3]
    float* s_mean = s_dc + kShThreads * 3;           // [TH*3]
    uint64_t* bar = reinterpret_cast<uint64_t*>(s_mean + kShThreads * 3);
    const int item0 = blockIdx.x * kShThreads;
    const int n_valid = min(kShThreads, N - item0);
    const int tid = threadIdx.x;
    if (n_valid == kShThreads) {
        if (tid == 0) {
            mbar_init(bar, 1);
            fence_proxy_async();
            const unsigned b_rest = kShThreads * R * 4, b_3 = kShThreads * 3 * 4;
            mbar_expect_tx(bar, b_rest + 2 * b_3);
            bulk_g2s(s_rest, rest + (size_t)item0 * R, b_rest, bar);
            bulk_g2s(s_dc, dc + (size_t)item0 * 3, b_3, bar);
            bulk_g2s(s_mean, means + (size_t)item0 * 3, b_3, bar);
        }
        __syncthreads();          // barrier init visible before anyone waits
        mbar_wait(bar, 0);
    } else {
        dense_copy_in<kShThreads>(s_rest, rest + (size_t)item0 * R, n_valid * R);
        dense_copy_in<kShThreads>(s_dc, dc + (size_t)item0 * 3, n_valid * 3);
        dense_copy_in<kShThreads>(s_mean, means + (size_t)item0 * 3, n_valid * 3);
        __syncthreads();
    }
    if (tid >= n_valid) return;
    float b[NB];
    float dx, dy, dz;
    load_dir(s_mean, tid, flags, viewmat, dx, dy, dz);
    sh_basis<DEG>(dx, dy, dz, b);
    float r = b[0] * s_dc[3 * tid], g = b[0] * s_dc[3 * tid + 1], bl = b[0] * s_dc[3 * tid + 2];
    const float* c = s_rest + tid * R;
#pragma unroll
    for (int k = 1; k < NB; ++k) {
        r = fmaf(b[k], c[3 * (k - 1)], r);
        g = fmaf(b[k], c[3 * (k - 1) + 1], g);
        bl = fmaf(b[k], c[3 * (k - 1) + 2], bl);
    }
    if (flags & TS_SH_OFFSET_CLAMP) {
        r += 0.5f; g += 0.5f; bl += 0.5f;
        unsigned m = (r >= 0.f ? 1u : 0u) | (g >= 0.f ? 2u : 0u) | (bl >= 0.f ? 4u : 0u);
        r = fmaxf(r, 0.f); g = fmaxf(g, 0.f); bl = fmaxf(bl, 0.f);
        if (clamp_mask) clamp_mask[item0 + tid] = (uint8_t)m;
    }
    float* o = colors + (size_t)out_stride * (item0 + tid);
    float d = ch3 ? __ldg(ch3 + item0 + tid) : 0.f;
    if ((out_stride & 3) == 0 && aligned_dev16(colors)) {
        *reinterpret_cast<float4*>(o) = make_float4(r, g, bl, d);
    } else {
        o[0] = r; o[1] = g; o[2] = bl;
        if (ch3) o[3] = d;
    }
}

template <int DEG>
__global__ void __launch_bounds__(kShThreads)
sh_bwd_bulk_kernel(int N, int K, const float* __restrict__ means, const float* __restrict__ viewmat,
                   const float* __restrict__ v_colors, int v_stride,
                   const uint8_t* __restrict__ clamp_mask, float* __restrict__ v_dc,
                   float* __restrict__ v_rest, int flags) {
    TS_DYN_SMEM(float, s_sh, 128);
    constexpr int NB = (DEG + 1) * (DEG + 1);
    const int R = (K - 1) * 3;
    float* s_rest = s_sh;                            // [TH][R] dense, becomes v_rest rows
    float* s_dc = s_rest + kShThreads * R;           // [TH*3]
    const int item0 = blockIdx.x * kShThreads;
    const int n_valid = min(kShThreads, N - item0);
    const int tid = threadIdx.x;
    if (tid < n_valid) {
        const int i = item0 + tid;
        float b[NB];
        float dx = __ldg(means + 3 * (size_t)i), dy = __ldg(means + 3 * (size_t)i + 1), dz = __ldg(means + 3 * (size_t)i + 2);
        if (flags & TS_SH_DIRS_FROM_MEANS) {
            dx -= __ldg(viewmat + 3); dy -= __ldg(viewmat + 7); dz -= __ldg(viewmat + 11);
        }
        sh_basis<DEG>(dx, dy, dz, b);
        float vr, vg, vb;
        const float* v = v_colors + (size_t)v_stride * i;
        if ((v_stride & 3) == 0 && aligned_dev16(v_colors)) {
            float4 t = __ldg(reinterpret_cast<const float4*>(v));
            vr = t.x; vg = t.y; vb = t.z;
        } else {
            vr = __ldg(v); vg = __ldg(v + 1); vb = __ldg(v + 2);
        }
        if (clamp_mask) {
            unsigned m = clamp_mask[i];
            if (!(m & 1u)) vr = 0.f;
            if (!(m & 2u)) vg = 0.f;
            if (!(m & 4u)) vb = 0.f;
        }
        s_dc[3 * tid] = b[0] * vr; s_dc[3 * tid + 1] = b[0] * vg; s_dc[3 * tid + 2] = b[0] * vb;
        float* c = s_rest + tid * R;
#pragma unroll
        for (int k = 1; k < NB; ++k) {
            c[3 * (k - 1)] = b[k] * vr;
            c[3 * (k - 1) + 1] = b[k] * vg;
            c[3 * (k - 1) + 2] = b[k] * vb;
        }
        for (int k = (NB - 1) * 3; k < R; ++k) c[k] = 0.f;   // bases above the active degree
    }
    if (n_valid == kShThreads) {
        fence_proxy_async();      // generic-proxy smem writes -> visible to the copy engine
        __syncthreads();
        if (tid == 0) {
            bulk_s2g(v_rest + (size_t)item0 * R, s_rest, kShThreads * R * 4);
            bulk_s2g(v_dc + (size_t)item0 * 3, s_dc, kShThreads * 3 * 4);
            bulk_commit();
            bulk_wait_read0();    // shared memory must outlive the reads
        }
    } else {
        __syncthreads();
        float* gr = v_rest + (size_t)item0 * R;
        for (int i = tid; i < n_valid * R; i += kShThreads) gr[i] = s_rest[i];
        float* gd = v_dc + (size_t)item0 * 3;
        for (int i = tid; i < n_valid * 3; i += kShThreads) gd[i] = s_dc[i];
    }
}

// Data-parallel shard backward (see project_bwd_views_kernel): sums SH-backward over the views
// whose packed colour cotangents (floats 8..10 of each 12-float packed row, clamp mask already
// applied by ts_dp_prepare) arrived for this rank's shard.  The view direction of view v is
// mean - cams[v][3,7,11] like TS_SH_DIRS_FROM_MEANS.  Rows are built in shared memory and leave
// with a TMA bulk store (full, aligned blocks) exactly like sh_bwd_bulk_kernel.
template <int DEG>
__global__ void __launch_bounds__(kShThreads)
sh_bwd_views_kernel(int n_views, int N, int K, const float* __restrict__ means,
                    const float* __restrict__ cams, const float* __restrict__ packed, size_t view_stride,
                    int row_stride, int col_off, float out_scale, float* __restrict__ v_dc,
                    float* __restrict__ v_rest) {
    TS_DYN_SMEM(float, s_sh, 128);
    constexpr int NB = (DEG + 1) * (DEG + 1);
    const int R = (K - 1) * 3;
    float* s_rest = s_sh;                            // [TH][R] dense, becomes v_rest rows
    float* s_dc = s_rest + kShThreads * R;           // [TH*3]
    const int item0 = blockIdx.x * kShThreads;
    const int n_valid = min(kShThreads, N - item0);
    const int tid = threadIdx.x;
    if (tid < n_valid) {
        const int i = item0 + tid;
        const float mx = __ldg(means + 3 * (size_t)i), my = __ldg(means + 3 * (size_t)i + 1), mz = __ldg(means + 3 * (size_t)i + 2);
        float acc[NB][3];
#pragma unroll
        for (int k = 0; k < NB; ++k) acc[k][0] = acc[k][1] = acc[k][2] = 0.f;
        for (int v = 0; v < n_views; ++v) {
            // colour cotangents of view v: floats col_off..col_off+2 of the row (packed record: 12 / 8,
            // one 128-bit load; colour rows of the peer exchange: 3 / 0)
            const float* src = packed + (size_t)v * view_stride + (size_t)row_stride * i + col_off;
            float4 t;
            if (((row_stride | col_off) & 3) == 0) {
                t = __ldg(reinterpret_cast<const float4*>(src));
            } else {
                t.x = __ldg(src); t.y = __ldg(src + 1); t.z = __ldg(src + 2); t.w = 0.f;
            }
            if (t.x == 0.f && t.y == 0.f && t.z == 0.f) continue;
            const float* cv = cams + (size_t)v * 32;
            float b[NB];
            sh_basis<DEG>(mx - __ldg(cv + 3), my - __ldg(cv + 7), mz - __ldg(cv + 11), b);
#pragma unroll
            for (int k = 0; k < NB; ++k) {
                acc[k][0] = fmaf(b[k], t.x, acc[k][0]);
                acc[k][1] = fmaf(b[k], t.y, acc[k][1]);
                acc[k][2] = fmaf(b[k], t.z, acc[k][2]);
            }
        }
        s_dc[3 * tid] = acc[0][0] * out_scale; s_dc[3 * tid + 1] = acc[0][1] * out_scale; s_dc[3 * tid + 2] = acc[0][2] * out_scale;
        float* c = s_rest + tid * R;
#pragma unroll
        for (int k = 1; k < NB; ++k) {
            c[3 * (k - 1)] = acc[k][0] * out_scale;
            c[3 * (k - 1) + 1] = acc[k][1] * out_scale;
            c[3 * (k - 1) + 2] = acc[k][2] * out_scale;
        }
        for (int k = (NB - 1) * 3; k < R; ++k) c[k] = 0.f;   // bases above the active degree
    }
    const bool bulk = n_valid == kShThreads && R > 0 && ((size_t)kShThreads * R * 4) % 16 == 0 &&
                      aligned_dev16(v_rest + (size_t)item0 * R) && aligned_dev16(v_dc + (size_t)item0 * 3);
    if (bulk) {
        fence_proxy_async();      // generic-proxy smem writes -> visible to the copy engine
        __syncthreads();
        if (tid == 0) {
            bulk_s2g(v_rest + (size_t)item0 * R, s_rest, kShThreads * R * 4);
            bulk_s2g(v_dc + (size_t)item0 * 3, s_dc, kShThreads * 3 * 4);
            bulk_commit();
            bulk_wait_read0();    // shared memory must outlive the reads
        }
    } else {
        __syncthreads();
        float* gr = v_rest + (size_t)item0 * R;
        for (int i = tid; i < n_valid * R; i += kShThreads) gr[i] = s_rest[i];
        float* gd = v_dc + (size_t)item0 * 3;
        for (int i = tid; i < n_valid * 3; i += kShThreads) gd[i] = s_dc[i];
    }
}

// The same sum over views for the peer-memory exchange (peer.cu): the colour cotangents arrive as
// dense 12-byte rows rgb[v][i][3] in THIS rank's buffer, so the block's rows of every view (and its
// means) are contiguous chunks: one elected thread pulls them with TMA bulk loads (n_views + 1 copies,
// one mbarrier), the 128 threads build the 192-byte SH gradient rows in shared memory and they leave
// with TMA bulk stores.  Every rank runs this for ALL Gaussians (the SH gradient is rebuilt from its
// rank-<=world factors instead of being shipped: half the NVLink bytes).
template <int DEG>
__global__ void __launch_bounds__(kShThreads)
sh_bwd_views_rgb_kernel(int n_views, int N, int K, const float* __restrict__ means,
                        const float* __restrict__ cams, const float* __restrict__ rgb, size_t view_stride,
                        float out_scale, float* __restrict__ v_dc, float* __restrict__ v_rest) {
    TS_DYN_SMEM(float, s_sh, 128);
    constexpr int NB = (DEG + 1) * (DEG + 1);
    constexpr int TH = kShThreads;
    const int R = (K - 1) * 3;
    float* s_rest = s_sh;                            // [TH][R] dense, becomes v_rest rows
    float* s_dc = s_rest + TH * R;                   // [TH*3]
    float* s_mean = s_dc + TH * 3;                   // [TH*3]
    float* s_rgb = s_mean + TH * 3;                  // [n_views][TH*3]
    uint64_t* bar = reinterpret_cast<uint64_t*>(s_rgb + (size_t)n_views * TH * 3);
    const int item0 = blockIdx.x * TH;
    const int n_valid = min(TH, N - item0);
    const int tid = threadIdx.x;
    const bool full = n_valid == TH && aligned_dev16(means) && aligned_dev16(rgb) && (view_stride & 3) == 0;
    if (full) {
        if (tid == 0) {
            mbar_init(bar, 1);
            fence_proxy_async();
            mbar_expect_tx(bar, (unsigned)((n_views + 1) * TH * 12));
            bulk_g2s(s_mean, means + (size_t)item0 * 3, TH * 12, bar);
            for (int v = 0; v < n_views; ++v)
                bulk_g2s(s_rgb + (size_t)v * TH * 3, rgb + (size_t)v * view_stride + (size_t)item0 * 3, TH * 12, bar);
        }
        __syncthreads();          // barrier init visible before anyone waits
        mbar_wait(bar, 0);
    } else {
        dense_copy_in<TH>(s_mean, means + (size_t)item0 * 3, n_valid * 3);
        for (int v = 0; v < n_views; ++v)
            dense_copy_in<TH>(s_rgb + (size_t)v * TH * 3, rgb + (size_t)v * view_stride + (size_t)item0 * 3, n_valid * 3);
        __syncthreads();
    }
    if (tid < n_valid) {
        const float mx = s_mean[3 * tid], my = s_mean[3 * tid + 1], mz = s_mean[3 * tid + 2];
        float acc[NB][3];
#pragma unroll
        for (int k = 0; k < NB; ++k) acc[k][0] = acc[k][1] = acc[k][2] = 0.f;
        for (int v = 0; v < n_views; ++v) {
            const float* t = s_rgb + (size_t)v * TH * 3 + 3 * tid;     // stride 3: bank-conflict free
            const float tx = t[0], ty = t[1], tz = t[2];
            if (tx == 0.f && ty == 0.f && tz == 0.f) continue;
            const float* cv = cams + (size_t)v * 32;
            float b[NB];
            sh_basis<DEG>(mx - __ldg(cv + 3), my - __ldg(cv + 7), mz - __ldg(cv + 11), b);
#pragma unroll
            for (int k = 0; k < NB; ++k) {
                acc[k][0] = fmaf(b[k], tx, acc[k][0]);
                acc[k][1] = fmaf(b[k], ty, acc[k][1]);
                acc[k][2] = fmaf(b[k], tz, acc[k][2]);
            }
        }
        s_dc[3 * tid] = acc[0][0] * out_scale; s_dc[3 * tid + 1] = acc[0][1] * out_scale; s_dc[3 * tid + 2] = acc[0][2] * out_scale;
        float* c = s_rest + tid * R;
#pragma unroll
        for (int k = 1; k < NB; ++k) {
            c[3 * (k - 1)] = acc[k][0] * out_scale;
            c[3 * (k - 1) + 1] = acc[k][1] * out_scale;
            c[3 * (k - 1) + 2] = acc[k][2] * out_scale;
        }
        for (int k = (NB - 1) * 3; k < R; ++k) c[k] = 0.f;   // bases above the active degree
    }
    const bool bulk = n_valid == TH && R > 0 && ((size_t)TH * R * 4) % 16 == 0 &&
                      aligned_dev16(v_rest + (size_t)item0 * R) && aligned_dev16(v_dc + (size_t)item0 * 3);
    if (bulk) {
        fence_proxy_async();      // generic-proxy smem writes -> visible to the copy engine
        __syncthreads();
        if (tid == 0) {
            bulk_s2g(v_rest + (size_t)item0 * R, s_rest, TH * R * 4);
            bulk_s2g(v_dc + (size_t)item0 * 3, s_dc, TH * 3 * 4);
            bulk_commit();
            bulk_wait_read0();    // shared memory must outlive the reads
        }
    } else {
        __syncthreads();
        float* gr = v_rest + (size_t)item0 * R;
        for (int i = tid; i < n_valid * R; i += TH) gr[i] = s_rest[i];
        float* gd = v_dc + (size_t)item0 * 3;
        for (int i = tid; i < n_valid * 3; i += TH) gd[i] = s_dc[i];
    }
}

static inline int sh_stride(int K) { int k3 = K * 3; return (k3 & 1) ? k3 : k3 + 1; }

}  // namespace ts

#ifndef TS_HOST_EMU
extern "C" {

int ts_sh_fwd(int N, int degree, int K, const float* dirs, const float* viewmat, const float* coeffs,
              const float* coeffs_rest, float* colors, int out_stride, const float* ch3,
              uint8_t* clamp_mask, int flags, ts_stream_t stream) {
    if (N < 0 || degree < 0 || degree > 4 || K < (degree + 1) * (degree + 1) || K > 25 || out_stride < 3)
        return TS_ERR_INVALID;
    if (N == 0) return TS_OK;
    if (!dirs || !coeffs || !colors) return TS_ERR_INVALID;
    if ((flags & TS_SH_DIRS_FROM_MEANS) && !viewmat) return TS_ERR_INVALID;
    if (!ts::aligned16(dirs) || !ts::aligned16(colors) || !ts::aligned16(coeffs) ||
        (coeffs_rest && !ts::aligned16(coeffs_rest)))
        return TS_ERR_ALIGN;
    int sstride = ts::sh_stride(K);
    size_t smem = sizeof(float) * (ts::kShThreads * 3 + 4 + (size_t)ts::kShThreads * sstride);
    int grid = (N + ts::kShThreads - 1) / ts::kShThreads;
    cudaStream_t st = (cudaStream_t)stream;
    // TMA path: (dc, rest) split, every stored band active, strided output (the fused pipeline)
    if (coeffs_rest && K == (degree + 1) * (degree + 1) && K > 1 && out_stride != 3 &&
        ((K - 1) * 3 * 4 * ts::kShThreads) % 16 == 0) {
        size_t bsmem = sizeof(float) * ts::kShThreads * ((K - 1) * 3 + 6) + 16;
#define TS_LAUNCH_SH_FWD_BULK(D) \
    ts::sh_fwd_bulk_kernel<D><<<grid, ts::kShThreads, bsmem, st>>>(N, K, dirs, viewmat, coeffs, coeffs_rest, colors, out_stride, ch3, clamp_mask, flags)
        switch (degree) {
            case 1: TS_LAUNCH_SH_FWD_BULK(1); break;
            case 2: TS_LAUNCH_SH_FWD_BULK(2); break;
            case 3: TS_LAUNCH_SH_FWD_BULK(3); break;
            default: TS_LAUNCH_SH_FWD_BULK(4); break;
        }
#undef TS_LAUNCH_SH_FWD_BULK
        TS_CHECK_LAUNCH("ts_sh_fwd/bulk");
        return TS_OK;
    }
#define TS_LAUNCH_SH_FWD(D) \
    ts::sh_fwd_kernel<D><<<grid, ts::kShThreads, smem, st>>>(N, K, dirs, viewmat, coeffs, coeffs_rest, colors, out_stride, ch3, clamp_mask, flags, sstride)
    switch (degree) {
        case 0: TS_LAUNCH_SH_FWD(0); break;
        case 1: TS_LAUNCH_SH_FWD(1); break;
        case 2: TS_LAUNCH_SH_FWD(2); break;
        case 3: TS_LAUNCH_SH_FWD(3); break;
        default: TS_LAUNCH_SH_FWD(4); break;
    }
#undef TS_LAUNCH_SH_FWD
    TS_CHECK_LAUNCH("ts_sh_fwd");
    return TS_OK;
}

int ts_sh_bwd(int N, int degree, int K, const float* dirs, const float* viewmat, const float* v_colors,
              int v_stride, const uint8_t* clamp_mask, float* v_coeffs, float* v_coeffs_rest, int flags,
              ts_stream_t stream) {
    if (N < 0 || degree < 0 || degree > 4 || K < (degree + 1) * (degree + 1) || K > 25 || v_stride < 3)
        return TS_ERR_INVALID;
    if (N == 0) return TS_OK;
    if (!dirs || !v_colors || !v_coeffs) return TS_ERR_INVALID;
    if ((flags & TS_SH_DIRS_FROM_MEANS) && !viewmat) return TS_ERR_INVALID;
    if (!ts::aligned16(dirs) || !ts::aligned16(v_colors) || !ts::aligned16(v_coeffs) ||
        (v_coeffs_rest && !ts::aligned16(v_coeffs_rest)))
        return TS_ERR_ALIGN;
    int sstride = ts::sh_stride(K);
    size_t smem = sizeof(float) * (2 * (ts::kShThreads * 3 + 4) + (size_t)ts::kShThreads * sstride);
    int grid = (N + ts::kShThreads - 1) / ts::kShThreads;
    cudaStream_t st = (cudaStream_t)stream;
    if (v_coeffs_rest && K > 1 && v_stride != 3 && ((K - 1) * 3 * 4 * ts::kShThreads) % 16 == 0) {
        size_t bsmem = sizeof(float) * ts::kShThreads * ((K - 1) * 3 + 3) + 16;
#define TS_LAUNCH_SH_BWD_BULK(D) \
    ts::sh_bwd_bulk_kernel<D><<<grid, ts::kShThreads, bsmem, st>>>(N, K, dirs, viewmat, v_colors, v_stride, clamp_mask, v_coeffs, v_coeffs_rest, flags)
        switch (degree) {
            case 0: TS_LAUNCH_SH_BWD_BULK(0); break;
            case 1: TS_LAUNCH_SH_BWD_BULK(1); break;
            case 2: TS_LAUNCH_SH_BWD_BULK(2); break;
            case 3: TS_LAUNCH_SH_BWD_BULK(3); break;
            default: TS_LAUNCH_SH_BWD_BULK(4); break;
        }
#undef TS_LAUNCH_SH_BWD_BULK
        TS_CHECK_LAUNCH("ts_sh_bwd/bulk");
        return TS_OK;
    }
#define TS_LAUNCH_SH_BWD(D) \
    ts::sh_bwd_kernel<D><<<grid, ts::kShThreads, smem, st>>>(N, K, dirs, viewmat, v_colors, v_stride, clamp_mask, v_coeffs, v_coeffs_rest, flags, sstride)
    switch (degree) {
        case 0: TS_LAUNCH_SH_BWD(0); break;
        case 1: TS_LAUNCH_SH_BWD(1); break;
        case 2: TS_LAUNCH_SH_BWD(2); break;
        case 3: TS_LAUNCH_SH_BWD(3); break;
        default: TS_LAUNCH_SH_BWD(4); break;
    }
#undef TS_LAUNCH_SH_BWD
    TS_CHECK_LAUNCH("ts_sh_bwd");
    return TS_OK;
}

static int launch_sh_bwd_views(int n_views, int N, int degree, int K, const float* means, const float* cams,
                               const float* rows, int64_t view_stride_floats, int row_stride, int col_off,
                               float out_scale, float* v_dc, float* v_rest, ts_stream_t stream) {
    if (n_views < 1 || N < 0 || degree < 0 || degree > 4 || K < (degree + 1) * (degree + 1) || K > 25 ||
        view_stride_floats < 0 || (view_stride_floats % 4) != 0 || row_stride < 3 || col_off < 0)
        return TS_ERR_INVALID;
    if (N == 0) return TS_OK;
    if (!means || !cams || !rows || !v_dc || (K > 1 && !v_rest)) return TS_ERR_INVALID;
    if (!ts::aligned16(rows)) return TS_ERR_ALIGN;
    int grid = (N + ts::kShThreads - 1) / ts::kShThreads;
    size_t bsmem = sizeof(float) * ts::kShThreads * ((K - 1) * 3 + 3) + 16;
    cudaStream_t st = (cudaStream_t)stream;
#define TS_LAUNCH_SH_VIEWS(D) \
    ts::sh_bwd_views_kernel<D><<<grid, ts::kShThreads, bsmem, st>>>(n_views, N, K, means, cams, rows, (size_t)view_stride_floats, row_stride, col_off, out_scale, v_dc, v_rest)
    switch (degree) {
        case 0: TS_LAUNCH_SH_VIEWS(0); break;
        case 1: TS_LAUNCH_SH_VIEWS(1); break;
        case 2: TS_LAUNCH_SH_VIEWS(2); break;
        case 3: TS_LAUNCH_SH_VIEWS(3); break;
        default: TS_LAUNCH_SH_VIEWS(4); break;
    }
#undef TS_LAUNCH_SH_VIEWS
    TS_CHECK_LAUNCH("ts_sh_bwd_views");
    return TS_OK;
}

int ts_sh_bwd_views(int n_views, int N, int degree, int K, const float* means, const float* cams,
                    const float* packed_grads, int64_t view_stride_floats, float out_scale, float* v_dc,
                    float* v_rest, ts_stream_t stream) {
    return launch_sh_bwd_views(n_views, N, degree, K, means, cams, packed_grads, view_stride_floats, 12, 8,
                               out_scale, v_dc, v_rest, stream);
}

int ts_sh_bwd_views_rgb(int n_views, int N, int degree, int K, const float* means, const float* cams,
                        const float* rgb_rows, int64_t view_stride_floats, float out_scale, float* v_dc,
                        float* v_rest, ts_stream_t stream) {
    if (n_views < 1 || n_views > 64 || N < 0 || degree < 0 || degree > 4 || K < (degree + 1) * (degree + 1) ||
        K > 25 || view_stride_floats < 0 || (view_stride_floats % 4) != 0)
        return TS_ERR_INVALID;
    if (N == 0) return TS_OK;
    if (!means || !cams || !rgb_rows || !v_dc || (K > 1 && !v_rest)) return TS_ERR_INVALID;
    if (!ts::aligned16(rgb_rows)) return TS_ERR_ALIGN;
    int grid = (N + ts::kShThreads - 1) / ts::kShThreads;
    size_t bsmem = sizeof(float) * ts::kShThreads * ((K - 1) * 3 + 6 + 3 * n_views) + 16;
    cudaStream_t st = (cudaStream_t)stream;
#define TS_LAUNCH_SH_VIEWS_RGB(D)                                                                             \
    do {                                                                                                      \
        TS_CHECK_CUDA(cudaFuncSetAttribute(ts::sh_bwd_views_rgb_kernel<D>,                                    \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsmem),          \
                      "ts_sh_bwd_views_rgb/attr");                                                            \
        ts::sh_bwd_views_rgb_kernel<D><<<grid, ts::kShThreads, bsmem, st>>>(                                  \
            n_views, N, K, means, cams, rgb_rows, (size_t)view_stride_floats, out_scale, v_dc, v_rest);      \
    } while (0)
    switch (degree) {
        case 0: TS_LAUNCH_SH_VIEWS_RGB(0); break;
        case 1: TS_LAUNCH_SH_VIEWS_RGB(1); break;
        case 2: TS_LAUNCH_SH_VIEWS_RGB(2); break;
        case 3: TS_LAUNCH_SH_VIEWS_RGB(3); break;
        default: TS_LAUNCH_SH_VIEWS_RGB(4); break;
    }
#undef TS_LAUNCH_SH_VIEWS_RGB
    TS_CHECK_LAUNCH("ts_sh_bwd_views_rgb");
    return TS_OK;
}

}  // extern "C"
#endif  // !TS_HOST_EMU
