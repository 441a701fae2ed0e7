// SURVEY.md 8(f)-4 — fused SSIM forward / backward for the training loss
// [REF scripts/train.py:60-62: loss_ssim = 1 - model.ssim(rendered[1,3,H,W], gt[1,3,H,W]);
//  REF tinysplat/splatting/model_gaussian.py:57: SSIM(data_range=1.0, size_average=True, channel=3)].
// 11-tap separable Gaussian window, 'valid' filtering.  One CTA per 16x16 output tile and
// (batch, channel): the 26x26 input patches of X and Y are staged in shared memory ONCE and all
// five filtered moments (X, Y, X^2, Y^2, XY) come from that one read — a conv2d formulation reads
// and writes the full-size image ~20 times.  Inputs are addressed through element strides, so the
// [H,W,3] rendered image is read in place (no permute/contiguous copy).  Forward also stores the
// three partial derivatives backward needs; backward is the transposed (full) filtering of those.
#include "ts_common.cuh"

namespace ts {

constexpr int kWin = 11;
constexpr int kHalo = kWin - 1;           // 10
constexpr int kST = 16;                   // output tile edge
constexpr int kSP = kST + kHalo;          // 26: input patch edge

struct SsimWin { float w[kWin]; };
struct Strides { long long b, c, h, w; };

__global__ void __launch_bounds__(kST * kST)
ssim_fwd_kernel(int C, int H, int W, const float* __restrict__ X, const float* __restrict__ Y, Strides sx,
                Strides sy, SsimWin win, float C1, float C2, float* __restrict__ ssim_sum,
                float* __restrict__ dmu, float* __restrict__ de11, float* __restrict__ de12) {
    __shared__ float s_x[kSP][kSP + 1], s_y[kSP][kSP + 1];
    __shared__ float s_h[5][kSP][kST + 1];       // horizontally filtered moments
    __shared__ float s_red[kST * kST / 32];
    const int Ho = H - kHalo, Wo = W - kHalo;
    const int bc = blockIdx.z, b = bc / C, c = bc - b * C;
    const int ox0 = blockIdx.x * kST, oy0 = blockIdx.y * kST;
    const int tid = threadIdx.y * kST + threadIdx.x;
    const float* xb = X + b * sx.b + c * sx.c;
    const float* yb = Y + b * sy.b + c * sy.c;
    for (int i = tid; i < kSP * kSP; i += kST * kST) {
        int py = i / kSP, px = i - py * kSP;
        int gy = oy0 + py, gx = ox0 + px;
        bool in = gy < H && gx < W;
        s_x[py][px] = in ? __ldg(xb + gy * sx.h + gx * sx.w) : 0.f;
        s_y[py][px] = in ? __ldg(yb + gy * sy.h + gx * sy.w) : 0.f;
    }
    __syncthreads();
    for (int i = tid; i < kSP * kST; i += kST * kST) {
        int py = i / kST, px = i - py * kST;
        float a = 0.f, bb = 0.f, aa = 0.f, b2 = 0.f, ab = 0.f;
#pragma unroll
        for (int k = 0; k < kWin; ++k) {
            float xv = s_x[py][px + k], yv = s_y[py][px + k], w = win.w[k];
            a = fmaf(w, xv, a); bb = fmaf(w, yv, bb);
            aa = fmaf(w, xv * xv, aa); b2 = fmaf(w, yv * yv, b2); ab = fmaf(w, xv * yv, ab);
        }
        s_h[0][py][px] = a; s_h[1][py][px] = bb; s_h[2][py][px] = aa; s_h[3][py][px] = b2; s_h[4][py][px] = ab;
    }
    __syncthreads();
    const int ox = ox0 + threadIdx.x, oy = oy0 + threadIdx.y;
    float val = 0.f;
    if (ox < Wo && oy < Ho) {
        float m1 = 0.f, m2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
        for (int k = 0; k < kWin; ++k) {
            float w = win.w[k];
            m1 = fmaf(w, s_h[0][threadIdx.y + k][threadIdx.x], m1);
            m2 = fmaf(w, s_h[1][threadIdx.y + k][threadIdx.x], m2);
            e11 = fmaf(w, s_h[2][threadIdx.y + k][threadIdx.x], e11);
            e22 = fmaf(w, s_h[3][threadIdx.y + k][threadIdx.x], e22);
            e12 = fmaf(w, s_h[4][threadIdx.y + k][threadIdx.x], e12);
        }
        float s1 = e11 - m1 * m1, s2 = e22 - m2 * m2, s12 = e12 - m1 * m2;
        float A1 = 2.f * m1 * m2 + C1, A2 = 2.f * s12 + C2;
        float B1 = m1 * m1 + m2 * m2 + C1, B2 = s1 + s2 + C2;
        float iB1 = 1.f / B1, iB2 = 1.f / B2;
        val = (A1 * iB1) * (A2 * iB2);
        if (dmu) {
            // S = A1 A2 / (B1 B2); X enters through m1, e11 (sigma1^2 = e11 - m1^2), e12 (sigma12 = e12 - m1 m2)
            float dS_ds1 = -val * iB2;
            float dS_ds12 = 2.f * A1 * iB1 * iB2;
            float dS_dm = 2.f * m2 * A2 * iB1 * iB2 - 2.f * m1 * val * iB1 - 2.f * m1 * dS_ds1 - m2 * dS_ds12;
            size_t o = ((size_t)bc * Ho + oy) * Wo + ox;
            dmu[o] = dS_dm; de11[o] = dS_ds1; de12[o] = dS_ds12;
        }
    }
    val = warp_sum(val);
    if ((tid & 31) == 0) s_red[tid >> 5] = val;
    __syncthreads();
    if (tid < 32) {
        float v = tid < kST * kST / 32 ? s_red[tid] : 0.f;
        v = warp_sum(v);
        if (tid == 0) atomicAdd(ssim_sum + bc, v);
    }
}

// d mean_S / d X(q) = (1/(Ho Wo)) * sum_o w(o -> q) [ dmu(o) + 2 X(q) de11(o) + Y(q) de12(o) ]
__global__ void __launch_bounds__(kST * kST)
ssim_bwd_kernel(int C, int H, int W, const float* __restrict__ X, const float* __restrict__ Y, Strides sx,
                Strides sy, SsimWin win, const float* __restrict__ dmu, const float* __restrict__ de11,
                const float* __restrict__ de12, const float* __restrict__ v_pc, float* __restrict__ v_X) {
    __shared__ float s_d[3][kSP][kSP + 1];
    __shared__ float s_h[3][kSP][kST + 1];
    const int Ho = H - kHalo, Wo = W - kHalo;
    const int bc = blockIdx.z, b = bc / C, c = bc - b * C;
    const int x0 = blockIdx.x * kST, y0 = blockIdx.y * kST;      // input-pixel tile
    const int tid = threadIdx.y * kST + threadIdx.x;
    // output positions o = q - k, k in 0..10  ->  patch origin (x0 - 10, y0 - 10)
    for (int i = tid; i < kSP * kSP; i += kST * kST) {
        int py = i / kSP, px = i - py * kSP;
        int oy = y0 - kHalo + py, ox = x0 - kHalo + px;
        bool in = oy >= 0 && oy < Ho && ox >= 0 && ox < Wo;
        size_t o = ((size_t)bc * Ho + oy) * Wo + ox;
        s_d[0][py][px] = in ? __ldg(dmu + o) : 0.f;
        s_d[1][py][px] = in ? __ldg(de11 + o) : 0.f;
        s_d[2][py][px] = in ? __ldg(de12 + o) : 0.f;
    }
    __syncthreads();
    // q.x = x0 + px uses patch columns px + (10 - k), weight w[k]  (o.x = q.x - k)
    for (int i = tid; i < kSP * kST; i += kST * kST) {
        int py = i / kST, px = i - py * kST;
        float a = 0.f, bb = 0.f, cc = 0.f;
#pragma unroll
        for (int k = 0; k < kWin; ++k) {
            float w = win.w[k];
            a = fmaf(w, s_d[0][py][px + kHalo - k], a);
            bb = fmaf(w, s_d[1][py][px + kHalo - k], bb);
            cc = fmaf(w, s_d[2][py][px + kHalo - k], cc);
        }
        s_h[0][py][px] = a; s_h[1][py][px] = bb; s_h[2][py][px] = cc;
    }
    __syncthreads();
    const int qx = x0 + threadIdx.x, qy = y0 + threadIdx.y;
    if (qx < W && qy < H) {
        float a = 0.f, bb = 0.f, cc = 0.f;
#pragma unroll
        for (int k = 0; k < kWin; ++k) {
            float w = win.w[k];
            a = fmaf(w, s_h[0][threadIdx.y + kHalo - k][threadIdx.x], a);
            bb = fmaf(w, s_h[1][threadIdx.y + kHalo - k][threadIdx.x], bb);
            cc = fmaf(w, s_h[2][threadIdx.y + kHalo - k][threadIdx.x], cc);
        }
        float xv = __ldg(X + b * sx.b + c * sx.c + qy * sx.h + qx * sx.w);
        float yv = __ldg(Y + b * sy.b + c * sy.c + qy * sy.h + qx * sy.w);
        float scale = __ldg(v_pc + bc) / ((float)Ho * (float)Wo);
        v_X[((size_t)bc * H + qy) * W + qx] = scale * (a + 2.f * xv * bb + yv * cc);
    }
}

}  // namespace ts

#ifndef TS_HOST_EMU
extern "C" {

int ts_ssim_fwd(int B, int C, int H, int W, const float* X, const int64_t* x_strides, const float* Y,
                const int64_t* y_strides, const float* win11_host, float C1, float C2, float* ssim_sum,
                float* dmu, float* de11, float* de12, ts_stream_t stream) {
    if (B <= 0 || C <= 0 || H <= ts::kHalo || W <= ts::kHalo) return TS_ERR_INVALID;
    if (!X || !Y || !x_strides || !y_strides || !win11_host || !ssim_sum) return TS_ERR_INVALID;
    if ((dmu || de11 || de12) && !(dmu && de11 && de12)) return TS_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    TS_CHECK_CUDA(cudaMemsetAsync(ssim_sum, 0, sizeof(float) * (size_t)B * C, st), "ts_ssim_fwd/memset");
    ts::SsimWin win;
    for (int k = 0; k < ts::kWin; ++k) win.w[k] = win11_host[k];
    ts::Strides sx{x_strides[0], x_strides[1], x_strides[2], x_strides[3]};
    ts::Strides sy{y_strides[0], y_strides[1], y_strides[2], y_strides[3]};
    int Ho = H - ts::kHalo, Wo = W - ts::kHalo;
    dim3 grid((Wo + ts::kST - 1) / ts::kST, (Ho + ts::kST - 1) / ts::kST, B * C), block(ts::kST, ts::kST);
    ts::ssim_fwd_kernel<<<grid, block, 0, st>>>(C, H, W, X, Y, sx, sy, win, C1, C2, ssim_sum, dmu, de11, de12);
    TS_CHECK_LAUNCH("ts_ssim_fwd");
    return TS_OK;
}

int ts_ssim_bwd(int B, int C, int H, int W, const float* X, const int64_t* x_strides, const float* Y,
                const int64_t* y_strides, const float* win11_host, const float* dmu, const float* de11,
                const float* de12, const float* v_per_channel, float* v_X, ts_stream_t stream) {
    if (B <= 0 || C <= 0 || H <= ts::kHalo || W <= ts::kHalo) return TS_ERR_INVALID;
    if (!X || !Y || !x_strides || !y_strides || !win11_host || !dmu || !de11 || !de12 || !v_per_channel || !v_X)
        return TS_ERR_INVALID;
    ts::SsimWin win;
    for (int k = 0; k < ts::kWin; ++k) win.w[k] = win11_host[k];
    ts::Strides sx{x_strides[0], x_strides[1], x_strides[2], x_strides[3]};
    ts::Strides sy{y_strides[0], y_strides[1], y_strides[2], y_strides[3]};
    dim3 grid((W + ts::kST - 1) / ts::kST, (H + ts::kST - 1) / ts::kST, B * C), block(ts::kST, ts::kST);
    ts::ssim_bwd_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(C, H, W, X, Y, sx, sy, win, dmu, de11, de12,
                                                                 v_per_channel, v_X);
    TS_CHECK_LAUNCH("ts_ssim_bwd");
    return TS_OK;
}

}  // extern "C"
#endif  // !TS_HOST_EMU
