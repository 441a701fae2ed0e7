// K1 / K6 — EWA projection forward and backward (streaming, HBM-bound; one thread per
// Gaussian, AoS [N,3]/[N,6] arrays staged through shared memory with 128-bit accesses).
// Replaces gsplat.project_gaussians fwd/bwd  [REF tinysplat/splatting/rasterize.py:32,64-73].
#include <cstdlib>
#include "ts_common.cuh"
#include "ts_sh_basis.cuh"
#include "ts_binning.cuh"
#include "ts_peer.cuh"

namespace ts {

constexpr int kProjThreads = 256;

struct ProjCam {
    float V[12];  // 3x4 row-major world->camera
    float P[16];  // 4x4 row-major full projection
};

__device__ __forceinline__ void load_cam(const float* __restrict__ viewmat,
                                         const float* __restrict__ projmat, ProjCam& c) {
#pragma unroll
    for (int i = 0; i < 12; ++i) c.V[i] = __ldg(viewmat + i);
#pragma unroll
    for (int i = 0; i < 16; ++i) c.P[i] = __ldg(projmat + i);
}

// The reference adapter feeds exp(log-scales) and quats/|quats| [REF rasterize.py:72-73]; the
// fused pipeline folds those activations (and their Jacobians) into the kernels.
__device__ __forceinline__ float apply_activations(int flags, float sc[3], float4& q) {
    if (flags & TS_PROJ_LOG_SCALES) {
#pragma unroll
        for (int k = 0; k < 3; ++k) sc[k] = expf(sc[k]);
    }
    float qn = 1.f;
    if (flags & TS_PROJ_RAW_QUATS) {
        qn = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
        float inv = 1.f / qn;
        q.x *= inv; q.y *= inv; q.z *= inv; q.w *= inv;
    }
    return qn;
}

// Everything forward computes that backward needs again (recomputed, not stored).
struct ProjState {
    float t[3];
    float R[9];
    float s[3];        // glob_scale * scale
    float Sig[6];      // cov3d upper triangle: 00 01 02 11 12 22
    float T[6];        // 2x3
    float U[6];        // T * Sigma, 2x3
    float a, b, c, det;
    float J00, J02, J11, J12, rz;
    float txc, tyc, limx, limy;
    bool clampx, clampy;
    float ph[4], rw;
    bool near_ok, det_ok, w_ok;
};

__device__ __forceinline__ void project_core(const ProjCam& cam, const float mu[3],
                                             const float sc[3], float gs, float4 q, float fx,
                                             float fy, int H, int W, float clip, ProjState& st) {
    const float* V = cam.V;
#pragma unroll
    for (int r = 0; r < 3; ++r)
        st.t[r] = V[4 * r + 0] * mu[0] + V[4 * r + 1] * mu[1] + V[4 * r + 2] * mu[2] + V[4 * r + 3];
    st.near_ok = st.t[2] > clip;
    float tz = st.near_ok ? st.t[2] : 1.f;

    quat_to_rotmat(q, st.R);
#pragma unroll
    for (int j = 0; j < 3; ++j) st.s[j] = gs * sc[j];
    float s2[3] = {st.s[0] * st.s[0], st.s[1] * st.s[1], st.s[2] * st.s[2]};
    const float* R = st.R;
    st.Sig[0] = R[0] * R[0] * s2[0] + R[1] * R[1] * s2[1] + R[2] * R[2] * s2[2];
    st.Sig[1] = R[0] * R[3] * s2[0] + R[1] * R[4] * s2[1] + R[2] * R[5] * s2[2];
    st.Sig[2] = R[0] * R[6] * s2[0] + R[1] * R[7] * s2[1] + R[2] * R[8] * s2[2];
    st.Sig[3] = R[3] * R[3] * s2[0] + R[4] * R[4] * s2[1] + R[5] * R[5] * s2[2];
    st.Sig[4] = R[3] * R[6] * s2[0] + R[4] * R[7] * s2[1] + R[5] * R[8] * s2[2];
    st.Sig[5] = R[6] * R[6] * s2[0] + R[7] * R[7] * s2[1] + R[8] * R[8] * s2[2];

    st.limx = kFovClamp * 0.5f * (float)W / fx;
    st.limy = kFovClamp * 0.5f * (float)H / fy;
    float rx = st.t[0] / tz, ry = st.t[1] / tz;
    st.clampx = (rx < -st.limx) || (rx > st.limx);
    st.clampy = (ry < -st.limy) || (ry > st.limy);
    st.txc = tz * fminf(st.limx, fmaxf(-st.limx, rx));
    st.tyc = tz * fminf(st.limy, fmaxf(-st.limy, ry));
    st.rz = 1.f / tz;
    float rz2 = st.rz * st.rz;
    st.J00 = fx * st.rz;
    st.J02 = -fx * st.txc * rz2;
    st.J11 = fy * st.rz;
    st.J12 = -fy * st.tyc * rz2;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        st.T[k] = st.J00 * V[k] + st.J02 * V[8 + k];
        st.T[3 + k] = st.J11 * V[4 + k] + st.J12 * V[8 + k];
    }
    const float* S = st.Sig;
    const float Sm[9] = {S[0], S[1], S[2], S[1], S[3], S[4], S[2], S[4], S[5]};
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int k = 0; k < 3; ++k)
            st.U[3 * r + k] = st.T[3 * r + 0] * Sm[k] + st.T[3 * r + 1] * Sm[3 + k] +
                              st.T[3 * r + 2] * Sm[6 + k];
    st.a = st.U[0] * st.T[0] + st.U[1] * st.T[1] + st.U[2] * st.T[2] + kCov2dBlur;
    st.b = st.U[0] * st.T[3] + st.U[1] * st.T[4] + st.U[2] * st.T[5];
    st.c = st.U[3] * st.T[3] + st.U[4] * st.T[4] + st.U[5] * st.T[5] + kCov2dBlur;
    st.det = st.a * st.c - st.b * st.b;
    st.det_ok = st.det != 0.f;

    const float* P = cam.P;
#pragma unroll
    for (int r = 0; r < 4; ++r)
        st.ph[r] = P[4 * r + 0] * mu[0] + P[4 * r + 1] * mu[1] + P[4 * r + 2] * mu[2] + P[4 * r + 3];
    float w = st.ph[3] + kWEps;
    st.w_ok = st.near_ok && (w != 0.f);
    st.rw = 1.f / (st.w_ok ? w : 1.f);
}

__global__ void __launch_bounds__(kProjThreads)
project_fwd_kernel(int N, const float* __restrict__ means, const float* __restrict__ scales,
                   float gs, const float4* __restrict__ quats,
                   const float* __restrict__ viewmat, const float* __restrict__ projmat,
                   float fx, float fy, float cx, float cy, int H, int W, int tbx, int tby,
                   float clip, int flags, float2* __restrict__ xys, float* __restrict__ depths,
                   int32_t* __restrict__ radii, float* __restrict__ conics,
                   int32_t* __restrict__ ntiles, float* __restrict__ cov3d,
                   const float* __restrict__ opacity, int cull, float4* __restrict__ recs,
                   int32_t* __restrict__ tile_counts) {
    constexpr int TH = kProjThreads;
    __shared__ __align__(16) float s_buf[TH * 9];
    const int item0 = blockIdx.x * TH;
    const int tid = threadIdx.x;
    block_load<3, TH>(means, s_buf, item0, N);
    block_load<3, TH>(scales, s_buf + 3 * TH, item0, N);
    __syncthreads();
    const int i = item0 + tid;
    const bool in = i < N;
    float mu[3] = {0.f, 0.f, 1.f}, sc[3] = {1.f, 1.f, 1.f};
    float4 q = make_float4(1.f, 0.f, 0.f, 0.f);
    if (in) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            mu[k] = s_buf[3 * tid + k];
            sc[k] = s_buf[3 * TH + 3 * tid + k];
        }
        q = __ldg(quats + i);
    }
    apply_activations(flags, sc, q);
    __syncthreads();  // inputs consumed; s_buf is reused for output staging

    ProjCam cam;
    load_cam(viewmat, projmat, cam);
    ProjState st;
    project_core(cam, mu, sc, gs, q, fx, fy, H, W, clip, st);

    float dets = st.det_ok ? st.det : 1.f;
    float inv = 1.f / dets;
    float con0 = st.c * inv, con1 = -st.b * inv, con2 = st.a * inv;
    float mid = 0.5f * (st.a + st.c);
    float disc = sqrtf(fmaxf(mid * mid - st.det, kEigFloor));
    float lam = fmaxf(mid + disc, mid - disc);
    float radius = ceilf(3.f * sqrtf(lam));
    if (!(radius >= 0.f)) radius = 0.f;  // NaN -> 0
    radius = fminf(radius, kMaxRadius);
    float px = 0.5f * (float)W * st.ph[0] * st.rw + cx - 0.5f;
    float py = 0.5f * (float)H * st.ph[1] * st.rw + cy - 0.5f;
    // NaN/inf-safe copies for the bbox only
    float bx = (px == px) ? fminf(fmaxf(px, -1e9f), 1e9f) : 0.f;
    float by = (py == py) ? fminf(fmaxf(py, -1e9f), 1e9f) : 0.f;
    int lox, loy, hix, hiy;
    tile_bbox(bx, by, radius, tbx, tby, lox, loy, hix, hiy);
    int area = (hix - lox) * (hiy - loy);
    // radius > 0 and det == det: a NaN covariance (zero-norm / NaN quaternion, NaN log-scale) passes
    // det != 0 and gives radius 0, whose tile box still has area 1; such a Gaussian is culled HERE,
    // exactly where ts_bin_emit (radii > 0) skips it — otherwise the fused tile count and the
    // emitted keys disagree and one key slot stays unwritten
    bool ok = in && st.near_ok && st.det_ok && st.w_ok && (area > 0) && (radius > 0.f) && (st.det == st.det);

    s_buf[3 * tid + 0] = ok ? con0 : 0.f;
    s_buf[3 * tid + 1] = ok ? con1 : 0.f;
    s_buf[3 * tid + 2] = ok ? con2 : 0.f;
#pragma unroll
    for (int k = 0; k < 6; ++k) s_buf[3 * TH + 6 * tid + k] = st.near_ok ? st.Sig[k] : 0.f;
    if (in) {
        xys[i] = ok ? make_float2(px, py) : make_float2(0.f, 0.f);
        depths[i] = ok ? st.t[2] : 0.f;
        radii[i] = ok ? (int32_t)radius : 0;
        if (ntiles) ntiles[i] = ok ? area : 0;
    }
    if (recs) {
        // fused pipeline: pack the geometry half of the raster record and count the tiles this
        // Gaussian's footprint can reach, here, while everything is still in registers
        // (replaces the separate ts_bin_count pass and its re-read of xys/conics/radii)
        int rlox = 0, rloy = 0, rhix = 0, rhiy = 0;
        if (ok) {
            float op = __ldg(opacity + i);
            if (flags & TS_PROJ_OPACITY_LOGIT) op = 1.f / (1.f + expf(-op));
            float hx, hy;
            footprint_extent(con0, con1, con2, op, cull, hx, hy);
            const float4 q0 = make_float4(px, py, hx, hy);
            recs[3 * (size_t)i] = q0;
            recs[3 * (size_t)i + 1] = make_float4(0.5f * kLog2e * con0, kLog2e * con1, 0.5f * kLog2e * con2, op);
            tile_rect(q0, radius, tbx, tby, cull, rlox, rloy, rhix, rhiy);
        }
        for_each_tile(rlox, rloy, rhix, rhiy, tbx, 0u, 0u,
                      [&](int tile, uint32_t, uint32_t) { atomicAdd(tile_counts + (size_t)tile * kCounterStride, 1); });
    }
    __syncthreads();
    if (conics) block_store<3, TH>(conics, s_buf, item0, N);
    if (cov3d) block_store<6, TH>(cov3d, s_buf + 3 * TH, item0, N);
}

// Backward of one Gaussian in ONE view: accumulates into vmu / vs / vq (pre-activation: the
// Jacobians of exp / normalise are applied once by the caller, they do not depend on the view)
// and vlogit.  Upstream cotangents: (vxy, vdep, vcon) and/or blend-backward's packed record
// (g0 = {S_x, S_y, S_xx, S_xy}, g1 = {S_yy, v_opacity, .., ..}, depth cotangent `vdep_packed`).
__device__ __forceinline__ void project_bwd_view(const ProjCam& cam, const float mu[3], const float sc[3],
                                                 float gs, float4 q, float fx, float fy, int H, int W,
                                                 float2 vxy, float vdep, float vcon[3], bool has_packed,
                                                 float4 g0, float4 g1, float vdep_packed, bool has_logit,
                                                 float opac_logit, float vmu[3], float vs[3], float4& vq,
                                                 float& vlogit, float2& vxy_total) {
    ProjState st;
    project_core(cam, mu, sc, gs, q, fx, fy, H, W, -3.0e38f, st);  // radii>0 => passed the near clip
    const float* V = cam.V;
    const float* P = cam.P;
    float inv = 1.f / st.det;
    float A = st.c * inv, B = -st.b * inv, C = st.a * inv;   // the conic
    if (has_packed) {
        // blend-backward's packed record: S-sums -> cotangents of xy and conic; the fused
        // pipeline's depth cotangent rides in colour channel 3; v_opacity in g1.y
        vxy.x += A * g0.x + B * g0.y;
        vxy.y += B * g0.x + C * g0.y;
        vcon[0] += 0.5f * g0.z;
        vcon[1] += g0.w;
        vcon[2] += 0.5f * g1.x;
        vdep += vdep_packed;
        if (has_logit) {
            float o = 1.f / (1.f + expf(-opac_logit));
            vlogit += g1.y * o * (1.f - o);
        }
    }
    vxy_total = vxy;
    // (1) pixel position
    float vndx = 0.5f * (float)W * vxy.x, vndy = 0.5f * (float)H * vxy.y;
    float vph0 = vndx * st.rw, vph1 = vndy * st.rw;
    float vph3 = -(vndx * st.ph[0] + vndy * st.ph[1]) * st.rw * st.rw;
#pragma unroll
    for (int k = 0; k < 3; ++k) vmu[k] += P[k] * vph0 + P[4 + k] * vph1 + P[12 + k] * vph3;
    // (2) depth
#pragma unroll
    for (int k = 0; k < 3; ++k) vmu[k] += V[8 + k] * vdep;
    // (3) conic = (c, -b, a) / det  ->  cov2d (a, b, c), det = a c - b^2.  Evaluated in COVARIANCE
    // space (through det) and not as -X v X with the conic X: for a needle-shaped Gaussian the
    // component of v_cov along the long axis is ~1e-7 of the products that form -X v X, is lost to
    // fp32 cancellation there, and is then multiplied by the long axis' variance on its way to the
    // scale gradient (measured: 30 % error on d/d log-scale of the long axis; with this form the
    // geometric gradients match the fp64 oracle like the oracle's own fp32 run does).
    float tdet = (st.c * vcon[0] - st.b * vcon[1] + st.a * vcon[2]) * inv;
    float vdet = -tdet * inv;
    float va = vcon[2] * inv + st.c * vdet;
    float vb = -vcon[1] * inv - 2.f * st.b * vdet;
    float vc = vcon[0] * inv + st.a * vdet;
    float hb = 0.5f * vb;
    // (4) cov2d = T Sigma T^T
    const float* T = st.T;
    const float* U = st.U;
    float vT[6];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        vT[k] = 2.f * (va * U[k] + hb * U[3 + k]);
        vT[3 + k] = 2.f * (hb * U[k] + vc * U[3 + k]);
    }
    float vSig[9];
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int k = 0; k < 3; ++k)
            vSig[3 * j + k] = T[j] * (va * T[k] + hb * T[3 + k]) + T[3 + j] * (hb * T[k] + vc * T[3 + k]);
    // v_J = v_T * Rv^T (only the four live entries)
    float vJ00 = vT[0] * V[0] + vT[1] * V[1] + vT[2] * V[2];
    float vJ02 = vT[0] * V[8] + vT[1] * V[9] + vT[2] * V[10];
    float vJ11 = vT[3] * V[4] + vT[4] * V[5] + vT[5] * V[6];
    float vJ12 = vT[3] * V[8] + vT[4] * V[9] + vT[5] * V[10];
    float rz = st.rz, rz2 = rz * rz, rz3 = rz2 * rz;
    float vtxc = -fx * rz2 * vJ02, vtyc = -fy * rz2 * vJ12;
    float vt[3];
    vt[2] = -fx * rz2 * vJ00 - fy * rz2 * vJ11 + 2.f * fx * st.txc * rz3 * vJ02 +
            2.f * fy * st.tyc * rz3 * vJ12;
    if (st.clampx) { vt[0] = 0.f; vt[2] += (st.txc * rz) * vtxc; } else vt[0] = vtxc;
    if (st.clampy) { vt[1] = 0.f; vt[2] += (st.tyc * rz) * vtyc; } else vt[1] = vtyc;
#pragma unroll
    for (int k = 0; k < 3; ++k) vmu[k] += V[k] * vt[0] + V[4 + k] * vt[1] + V[8 + k] * vt[2];
    // (5) Sigma = M M^T, M = R diag(s)
    const float* R = st.R;
    float vR[9];
#pragma unroll
    for (int ii = 0; ii < 3; ++ii)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            float vM = 2.f * (vSig[3 * ii + 0] * R[0 + j] * st.s[j] + vSig[3 * ii + 1] * R[3 + j] * st.s[j] +
                              vSig[3 * ii + 2] * R[6 + j] * st.s[j]);
            vR[3 * ii + j] = vM * st.s[j];
            vs[j] += gs * R[3 * ii + j] * vM;
        }
    float w = q.x, x = q.y, y = q.z, z = q.w;
    vq.x += 2.f * (-z * vR[1] + y * vR[2] + z * vR[3] - x * vR[5] - y * vR[6] + x * vR[7]);
    vq.y += 2.f * (y * vR[1] + z * vR[2] + y * vR[3] - 2.f * x * vR[4] - w * vR[5] + z * vR[6] +
                  w * vR[7] - 2.f * x * vR[8]);
    vq.z += 2.f * (-2.f * y * vR[0] + x * vR[1] + w * vR[2] + x * vR[3] + z * vR[5] - w * vR[6] +
                  z * vR[7] - 2.f * y * vR[8]);
    vq.w += 2.f * (-2.f * z * vR[0] - w * vR[1] + x * vR[2] + w * vR[3] - 2.f * z * vR[4] +
                  y * vR[5] + x * vR[6] + y * vR[7]);
}

// Jacobians of the activations the fused pipeline folds in (exp of log-scales, q = r/|r|): linear
// maps that depend on the Gaussian only, applied once after all views have been accumulated.
__device__ __forceinline__ void project_bwd_activations(int flags, const float sc[3], float4 q, float qn,
                                                        float vs[3], float4& vq) {
    if (flags & TS_PROJ_LOG_SCALES) {
#pragma unroll
        for (int k = 0; k < 3; ++k) vs[k] *= sc[k];           // d exp(l)/dl = exp(l)
    }
    if (flags & TS_PROJ_RAW_QUATS) {                           // q = r/|r|
        float d = vq.x * q.x + vq.y * q.y + vq.z * q.z + vq.w * q.w;
        float iq = 1.f / qn;
        vq.x = (vq.x - d * q.x) * iq; vq.y = (vq.y - d * q.y) * iq;
        vq.z = (vq.z - d * q.z) * iq; vq.w = (vq.w - d * q.w) * iq;
    }
}

__global__ void __launch_bounds__(kProjThreads)
project_bwd_kernel(int N, const float* __restrict__ means, const float* __restrict__ scales,
                   float gs, const float4* __restrict__ quats,
                   const float* __restrict__ viewmat, const float* __restrict__ projmat,
                   float fx, float fy, float cx, float cy, int H, int W, int flags,
                   const int32_t* __restrict__ radii, const float2* __restrict__ v_xys,
                   const float* __restrict__ v_depths, const float* __restrict__ v_conics,
                   const float4* __restrict__ packed, const float* __restrict__ opac_logits,
                   float* __restrict__ v_means, float* __restrict__ v_scales,
                   float4* __restrict__ v_quats, float* __restrict__ v_opac_logits,
                   float2* __restrict__ v_xys_out) {
    constexpr int TH = kProjThreads;
    __shared__ __align__(16) float s_buf[TH * 9];
    const int item0 = blockIdx.x * TH;
    const int tid = threadIdx.x;
    block_load<3, TH>(means, s_buf, item0, N);
    block_load<3, TH>(scales, s_buf + 3 * TH, item0, N);
    if (v_conics) block_load<3, TH>(v_conics, s_buf + 6 * TH, item0, N);
    __syncthreads();
    const int i = item0 + tid;
    const bool in = i < N;
    float mu[3] = {0.f, 0.f, 1.f}, sc[3] = {1.f, 1.f, 1.f}, vcon[3] = {0.f, 0.f, 0.f};
    float4 q = make_float4(1.f, 0.f, 0.f, 0.f);
    float2 vxy = make_float2(0.f, 0.f);
    float vdep = 0.f;
    bool ok = false;
    if (in) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            mu[k] = s_buf[3 * tid + k];
            sc[k] = s_buf[3 * TH + 3 * tid + k];
            if (v_conics) vcon[k] = s_buf[6 * TH + 3 * tid + k];
        }
        q = __ldg(quats + i);
        ok = __ldg(radii + i) > 0;
        if (v_xys) vxy = __ldg(v_xys + i);
        if (v_depths) vdep = __ldg(v_depths + i);
    }
    const float qn = apply_activations(flags, sc, q);
    __syncthreads();

    float vmu[3] = {0.f, 0.f, 0.f}, vs[3] = {0.f, 0.f, 0.f};
    float4 vq = make_float4(0.f, 0.f, 0.f, 0.f);
    float vlogit = 0.f;
    if (ok) {
        ProjCam cam;
        load_cam(viewmat, projmat, cam);
        float4 g0 = make_float4(0.f, 0.f, 0.f, 0.f), g1 = g0;
        float vdep_packed = 0.f, logit = 0.f;
        if (packed) {
            g0 = __ldg(packed + 3 * (size_t)i);
            g1 = __ldg(packed + 3 * (size_t)i + 1);
            if (flags & TS_PROJ_DEPTH_CH3) vdep_packed = __ldg(reinterpret_cast<const float*>(packed) + 12 * (size_t)i + 11);
            if (opac_logits) logit = __ldg(opac_logits + i);
        }
        project_bwd_view(cam, mu, sc, gs, q, fx, fy, H, W, vxy, vdep, vcon, packed != nullptr, g0, g1,
                         vdep_packed, opac_logits != nullptr, logit, vmu, vs, vq, vlogit, vxy);
        project_bwd_activations(flags, sc, q, qn, vs, vq);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        s_buf[3 * tid + k] = vmu[k];
        s_buf[3 * TH + 3 * tid + k] = vs[k];
    }
    if (in) {
        v_quats[i] = vq;
        if (v_opac_logits) v_opac_logits[i] = vlogit;
        if (v_xys_out) v_xys_out[i] = ok ? vxy : make_float2(0.f, 0.f);
    }
    __syncthreads();
    block_store<3, TH>(v_means, s_buf, item0, N);
    block_store<3, TH>(v_scales, s_buf + 3 * TH, item0, N);
}

// ---- K6 + K7 in one kernel (fused pipeline, single GPU) ------------------------------------------
// Projection-backward is issue-bound (74 registers, ~1000 instructions per Gaussian), SH-backward is
// DRAM-bound (192 B of coefficient gradient written per Gaussian).  As two kernels on two streams they
// take the SUM of their times: projection-backward reaches the SMs first and its 3 CTAs per SM leave
// room for one SH CTA (a priority stream or an occupancy cap did not help: profiles/README.md).  Here
// every CTA does both for its 256 Gaussians: the EWA algebra, then the SH rows built in shared memory
// leave as TMA bulk stores while other warps still compute — the copy engine and the issue slots overlap
// inside the kernel instead of depending on the block scheduler.
template <int DEG>
__global__ void __launch_bounds__(kProjThreads)
project_sh_bwd_kernel(int N, int K, const float* __restrict__ means, const float* __restrict__ scales,
                      float gs, const float4* __restrict__ quats, const float* __restrict__ viewmat,
                      const float* __restrict__ projmat, float fx, float fy, float cx, float cy, int H, int W,
                      int flags, const int32_t* __restrict__ radii, const float4* __restrict__ packed,
                      const float* __restrict__ opac_logits, const uint8_t* __restrict__ clamp_mask,
                      float* __restrict__ v_means, float* __restrict__ v_scales, float4* __restrict__ v_quats,
                      float* __restrict__ v_opac_logits, float2* __restrict__ v_xys_out,
                      float* __restrict__ v_dc, float* __restrict__ v_rest) {
    constexpr int TH = kProjThreads;
    constexpr int NB = (DEG + 1) * (DEG + 1);
    TS_DYN_SMEM(float, s_dyn, 128);
    const int R = (K - 1) * 3;
    float* s_rest = s_dyn;                       // [TH][R] dense, becomes v_rest rows
    float* s_dc = s_rest + TH * R;               // [TH*3]
    float* s_buf = s_dc + TH * 3;                // [TH*6]: means | scales in, their gradients out
    const int item0 = blockIdx.x * TH;
    const int tid = threadIdx.x;
    const int n_valid = min(TH, N - item0);
    block_load<3, TH>(means, s_buf, item0, N);
    block_load<3, TH>(scales, s_buf + 3 * TH, item0, N);
    __syncthreads();
    const int i = item0 + tid;
    const bool in = i < N;
    float mu[3] = {0.f, 0.f, 1.f}, sc[3] = {1.f, 1.f, 1.f}, vcon[3] = {0.f, 0.f, 0.f};
    float4 q = make_float4(1.f, 0.f, 0.f, 0.f);
    float2 vxy = make_float2(0.f, 0.f);
    bool ok = false;
    if (in) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            mu[k] = s_buf[3 * tid + k];
            sc[k] = s_buf[3 * TH + 3 * tid + k];
        }
        q = __ldg(quats + i);
        ok = __ldg(radii + i) > 0;
    }
    const float qn = apply_activations(flags, sc, q);
    __syncthreads();

    // ---- K7: the SH rows of this Gaussian (colour cotangent = floats 8..10 of its packed row; a culled
    // Gaussian's row is all zero, so its SH rows are zero as in sh_bwd_bulk_kernel)
    if (in) {
        float4 g2 = __ldg(packed + 3 * (size_t)i + 2);
        if (clamp_mask) {                        // SH clamp(rgb + 0.5, min=0) [REF rasterize.py:39]
            const unsigned m = clamp_mask[i];
            if (!(m & 1u)) g2.x = 0.f;
            if (!(m & 2u)) g2.y = 0.f;
            if (!(m & 4u)) g2.z = 0.f;
        }
        float b[NB];
        // view direction as the reference adapter forms it: mean - view-matrix translation column
        sh_basis<DEG>(mu[0] - __ldg(viewmat + 3), mu[1] - __ldg(viewmat + 7), mu[2] - __ldg(viewmat + 11), b);
        s_dc[3 * tid] = b[0] * g2.x; s_dc[3 * tid + 1] = b[0] * g2.y; s_dc[3 * tid + 2] = b[0] * g2.z;
        float* c = s_rest + tid * R;
#pragma unroll
        for (int k = 1; k < NB; ++k) {
            c[3 * (k - 1)] = b[k] * g2.x;
            c[3 * (k - 1) + 1] = b[k] * g2.y;
            c[3 * (k - 1) + 2] = b[k] * g2.z;
        }
        for (int k = (NB - 1) * 3; k < R; ++k) c[k] = 0.f;   // bases above the active degree
    }
    const bool bulk = n_valid == TH && R > 0 && ((size_t)TH * R * 4) % 16 == 0 &&
                      aligned_dev16(v_rest + (size_t)item0 * R) && aligned_dev16(v_dc + (size_t)item0 * 3);
    if (bulk) {
        fence_proxy_async();          // generic-proxy smem writes -> visible to the copy engine
        __syncthreads();
        if (tid == 0) {               // the copies run while the block does the EWA algebra below
            bulk_s2g(v_rest + (size_t)item0 * R, s_rest, TH * R * 4);
            bulk_s2g(v_dc + (size_t)item0 * 3, s_dc, TH * 3 * 4);
            bulk_commit();
        }
    } else {
        __syncthreads();
        float* gr = v_rest + (size_t)item0 * R;
        for (int k = tid; k < n_valid * R; k += TH) gr[k] = s_rest[k];
        float* gd = v_dc + (size_t)item0 * 3;
        for (int k = tid; k < n_valid * 3; k += TH) gd[k] = s_dc[k];
    }

    // ---- K6: projection-backward of this Gaussian
    float vmu[3] = {0.f, 0.f, 0.f}, vs[3] = {0.f, 0.f, 0.f};
    float4 vq = make_float4(0.f, 0.f, 0.f, 0.f);
    float vlogit = 0.f;
    if (ok) {
        ProjCam cam;
        load_cam(viewmat, projmat, cam);
        const float4 g0 = __ldg(packed + 3 * (size_t)i);
        const float4 g1 = __ldg(packed + 3 * (size_t)i + 1);
        float vdep_packed = 0.f, logit = 0.f;
        if (flags & TS_PROJ_DEPTH_CH3) vdep_packed = __ldg(reinterpret_cast<const float*>(packed) + 12 * (size_t)i + 11);
        if (opac_logits) logit = __ldg(opac_logits + i);
        project_bwd_view(cam, mu, sc, gs, q, fx, fy, H, W, vxy, 0.f, vcon, true, g0, g1, vdep_packed,
                         opac_logits != nullptr, logit, vmu, vs, vq, vlogit, vxy);
        project_bwd_activations(flags, sc, q, qn, vs, vq);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        s_buf[3 * tid + k] = vmu[k];
        s_buf[3 * TH + 3 * tid + k] = vs[k];
    }
    if (in) {
        v_quats[i] = vq;
        if (v_opac_logits) v_opac_logits[i] = vlogit;
        if (v_xys_out) v_xys_out[i] = ok ? vxy : make_float2(0.f, 0.f);
    }
    __syncthreads();
    block_store<3, TH>(v_means, s_buf, item0, N);
    block_store<3, TH>(v_scales, s_buf + 3 * TH, item0, N);
    if (bulk && tid == 0) bulk_wait_read0();     // shared memory must outlive the copy engine's reads
}

// ---- data-parallel shard backward ------------------------------------------------
// One rank owns a SHARD of the Gaussians and receives, from every rank/view v, blend-backward's
// packed gradient rows of that shard (packed[v * view_stride + 3 i .. +2], already cleaned by
// ts_dp_prepare: all-zero rows for Gaussians culled in view v).  This kernel runs projection-
// backward for every view with that view's camera (cams[v]: 3x4 view | 4x4 full projection | fx,
// fy) and writes the SUM over views times out_scale: the exchanged payload is the 48-byte packed
// record per (view, Gaussian) instead of an all-reduce over the finished 236-byte gradients.
constexpr int kCamFloats = kCamRowFloats;

// COMPACT: rows are the 32-byte geometry rows of the peer exchange ({S_x, S_y, S_xx, S_xy},
// {S_yy, v_opacity, v_depth, 0}; peer.cu) instead of the full 48-byte packed record.
// The finished shard gradients are stored to n_dst destinations (the same rows of every rank's
// gradient buffer, reached through NVLink peer mappings: the all-gather of the exchange happens
// in this epilogue, tile by tile, instead of as a separate collective).
template <bool COMPACT>
__global__ void __launch_bounds__(kProjThreads)
project_bwd_views_kernel(int n_views, int N, const float* __restrict__ means,
                         const float* __restrict__ scales, float gs, const float4* __restrict__ quats,
                         const float* __restrict__ cams, int H, int W, int flags,
                         const float4* __restrict__ packed, size_t view_stride4,
                         const float* __restrict__ opac_logits, float out_scale, int n_dst, int first_dst,
                         PeerPtrs d_means, PeerPtrs d_scales, PeerPtrs d_quats, PeerPtrs d_logits) {
    constexpr int TH = kProjThreads;
    constexpr int ROW4 = COMPACT ? 2 : 3;
    __shared__ __align__(128) float s_buf[TH * 11];     // means 3 | scales 3 | quats 4 | logit 1 (output staging)
    const int item0 = blockIdx.x * TH;
    const int tid = threadIdx.x;
    block_load<3, TH>(means, s_buf, item0, N);
    block_load<3, TH>(scales, s_buf + 3 * TH, item0, N);
    __syncthreads();
    const int i = item0 + tid;
    const bool in = i < N;
    float mu[3] = {0.f, 0.f, 1.f}, sc[3] = {1.f, 1.f, 1.f};
    float4 q = make_float4(1.f, 0.f, 0.f, 0.f);
    float logit = 0.f;
    if (in) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            mu[k] = s_buf[3 * tid + k];
            sc[k] = s_buf[3 * TH + 3 * tid + k];
        }
        q = __ldg(quats + i);
        if (opac_logits) logit = __ldg(opac_logits + i);
    }
    const float qn = apply_activations(flags, sc, q);
    __syncthreads();

    float vmu[3] = {0.f, 0.f, 0.f}, vs[3] = {0.f, 0.f, 0.f};
    float4 vq = make_float4(0.f, 0.f, 0.f, 0.f);
    float vlogit = 0.f;
    if (in) {
        for (int v = 0; v < n_views; ++v) {
            const float4* row = packed + (size_t)v * view_stride4 + ROW4 * (size_t)i;
            const float4 g0 = __ldg(row);
            float4 g1 = __ldg(row + 1);
            float vdep_packed;
            if (COMPACT) {
                vdep_packed = g1.z;
            } else {
                vdep_packed = (flags & TS_PROJ_DEPTH_CH3) ? __ldg(reinterpret_cast<const float*>(row) + 11) : 0.f;
            }
            // culled in this view (the row was zeroed) or simply untouched: nothing to
            // add, and the projection of a culled Gaussian must not be evaluated (0 * inf)
            if (g0.x == 0.f && g0.y == 0.f && g0.z == 0.f && g0.w == 0.f && g1.x == 0.f && g1.y == 0.f &&
                vdep_packed == 0.f)
                continue;
            const float* cv = cams + (size_t)v * kCamFloats;
            ProjCam cam;
            load_cam(cv, cv + 12, cam);
            float vcon[3] = {0.f, 0.f, 0.f};
            float2 vxy_unused;
            project_bwd_view(cam, mu, sc, gs, q, __ldg(cv + 28), __ldg(cv + 29), H, W, make_float2(0.f, 0.f), 0.f,
                             vcon, true, g0, g1, vdep_packed, opac_logits != nullptr, logit, vmu, vs, vq, vlogit,
                             vxy_unused);
        }
        project_bwd_activations(flags, sc, q, qn, vs, vq);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        s_buf[3 * tid + k] = vmu[k] * out_scale;
        s_buf[3 * TH + 3 * tid + k] = vs[k] * out_scale;
    }
    const float4 vq_out = make_float4(vq.x * out_scale, vq.y * out_scale, vq.z * out_scale, vq.w * out_scale);
    const float vl_out = vlogit * out_scale;
    reinterpret_cast<float4*>(s_buf + 6 * TH)[tid] = vq_out;
    s_buf[10 * TH + tid] = vl_out;
    // Full blocks leave as TMA bulk stores (3 + 3 + 4 + 1 KB per destination, one issuing thread):
    // with n_dst > 1 the destinations are peer-mapped buffers and the copy engine drives NVLink.
    bool bulk = (N - item0) >= TH;
    for (int d = 0; d < n_dst && bulk; ++d)
        bulk = aligned_dev16(d_means.p[d]) && aligned_dev16(d_scales.p[d]) && aligned_dev16(d_quats.p[d]) &&
               (!d_logits.p[d] || aligned_dev16(d_logits.p[d]));
    if (bulk) {
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            for (int d = 0; d < n_dst; ++d) {
                const int r = (first_dst + d) % n_dst;      // start at the neighbour: spreads the NVLink traffic
                bulk_s2g(reinterpret_cast<float*>(d_means.p[r]) + (size_t)item0 * 3, s_buf, TH * 12);
                bulk_s2g(reinterpret_cast<float*>(d_scales.p[r]) + (size_t)item0 * 3, s_buf + 3 * TH, TH * 12);
                bulk_s2g(reinterpret_cast<float*>(d_quats.p[r]) + (size_t)item0 * 4, s_buf + 6 * TH, TH * 16);
                if (d_logits.p[r]) bulk_s2g(reinterpret_cast<float*>(d_logits.p[r]) + item0, s_buf + 10 * TH, TH * 4);
            }
            bulk_commit();
            bulk_wait0();
        }
        return;
    }
    __syncthreads();
    for (int d = 0; d < n_dst; ++d) {
        const int r = (first_dst + d) % n_dst;
        if (in) {
            reinterpret_cast<float4*>(d_quats.p[r])[i] = vq_out;
            if (d_logits.p[r]) reinterpret_cast<float*>(d_logits.p[r])[i] = vl_out;
        }
        block_store<3, TH>(reinterpret_cast<float*>(d_means.p[r]), s_buf, item0, N);
        block_store<3, TH>(reinterpret_cast<float*>(d_scales.p[r]), s_buf + 3 * TH, item0, N);
    }
}

}  // namespace ts

#ifndef TS_HOST_EMU
extern "C" {

int ts_project_fwd(int N, const float* means3d, const float* scales, float glob_scale,
                   const float* quats, const float* viewmat, const float* projmat, float fx,
                   float fy, float cx, float cy, int img_height, int img_width, int tiles_x,
                   int tiles_y, float clip_thresh, int flags, float* xys, float* depths,
                   int32_t* radii, float* conics, int32_t* num_tiles_hit, float* cov3d,
                   const float* opacity, int cull_mode, float* recs, int32_t* tile_counts,
                   ts_stream_t stream) {
    if (N < 0 || img_height <= 0 || img_width <= 0 || tiles_x <= 0 || tiles_y <= 0) return TS_ERR_INVALID;
    // fused pack+count mode is selected by tile_counts (recs may legitimately be NULL when N == 0)
    if (tile_counts && N > 0 && (!opacity || !recs)) return TS_ERR_INVALID;
    if (!tile_counts && recs) return TS_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    if (tile_counts)
        TS_CHECK_CUDA(cudaMemsetAsync(tile_counts, 0, sizeof(int32_t) * ts::kCounterStride * (size_t)tiles_x * tiles_y, st),
                      "ts_project_fwd/memset");
    if (N == 0) return TS_OK;
    if (!means3d || !scales || !quats || !viewmat || !projmat || !xys || !depths || !radii)
        return TS_ERR_INVALID;
    if (!recs && (!conics || !num_tiles_hit || !cov3d)) return TS_ERR_INVALID;
    if (!ts::aligned16(means3d) || !ts::aligned16(scales) || !ts::aligned16(quats) ||
        !ts::aligned16(xys) || (conics && !ts::aligned16(conics)) || (cov3d && !ts::aligned16(cov3d)) ||
        (recs && !ts::aligned16(recs)))
        return TS_ERR_ALIGN;
    int grid = (N + ts::kProjThreads - 1) / ts::kProjThreads;
    ts::project_fwd_kernel<<<grid, ts::kProjThreads, 0, st>>>(
        N, means3d, scales, glob_scale, (const float4*)quats, viewmat, projmat, fx, fy, cx, cy,
        img_height, img_width, tiles_x, tiles_y, clip_thresh, flags, (float2*)xys, depths, radii,
        conics, num_tiles_hit, cov3d, opacity, cull_mode, (float4*)recs, tile_counts);
    TS_CHECK_LAUNCH("ts_project_fwd");
    return TS_OK;
}

int ts_project_bwd(int N, const float* means3d, const float* scales, float glob_scale,
                   const float* quats, const float* viewmat, const float* projmat, float fx,
                   float fy, float cx, float cy, int img_height, int img_width, int flags,
                   const int32_t* radii, const float* v_xys, const float* v_depths,
                   const float* v_conics, const float* packed_grads, const float* opacity_logits,
                   float* v_means3d, float* v_scales, float* v_quats, float* v_opacity_logits,
                   float* v_xys_out, ts_stream_t stream) {
    if (N < 0 || img_height <= 0 || img_width <= 0) return TS_ERR_INVALID;
    if (N == 0) return TS_OK;
    if (!means3d || !scales || !quats || !viewmat || !projmat || !radii || !v_means3d || !v_scales ||
        !v_quats)
        return TS_ERR_INVALID;
    if (!ts::aligned16(means3d) || !ts::aligned16(scales) || !ts::aligned16(quats) ||
        (v_conics && !ts::aligned16(v_conics)) || (packed_grads && !ts::aligned16(packed_grads)) ||
        !ts::aligned16(v_means3d) || !ts::aligned16(v_scales) || !ts::aligned16(v_quats) ||
        (v_xys && (reinterpret_cast<uintptr_t>(v_xys) & 7u)) ||
        (v_xys_out && (reinterpret_cast<uintptr_t>(v_xys_out) & 7u)))
        return TS_ERR_ALIGN;
    int grid = (N + ts::kProjThreads - 1) / ts::kProjThreads;
    ts::project_bwd_kernel<<<grid, ts::kProjThreads, 0, (cudaStream_t)stream>>>(
        N, means3d, scales, glob_scale, (const float4*)quats, viewmat, projmat, fx, fy, cx, cy,
        img_height, img_width, flags, radii, (const float2*)v_xys, v_depths, v_conics,
        (const float4*)packed_grads, opacity_logits, v_means3d, v_scales, (float4*)v_quats,
        v_opacity_logits, (float2*)v_xys_out);
    TS_CHECK_LAUNCH("ts_project_bwd");
    return TS_OK;
}

int ts_project_sh_bwd(int N, int degree, int K, const float* means3d, const float* scales, float glob_scale,
                      const float* quats, const float* viewmat, const float* projmat, float fx, float fy,
                      float cx, float cy, int img_height, int img_width, int flags, const int32_t* radii,
                      const float* packed_grads, const float* opacity_logits, const uint8_t* clamp_mask,
                      float* v_means3d, float* v_scales, float* v_quats, float* v_opacity_logits,
                      float* v_xys_out, float* v_dc, float* v_rest, ts_stream_t stream) {
    if (N < 0 || img_height <= 0 || img_width <= 0 || degree < 0 || degree > 4 || K < (degree + 1) * (degree + 1) || K > 25)
        return TS_ERR_INVALID;
    if (N == 0) return TS_OK;
    if (!means3d || !scales || !quats || !viewmat || !projmat || !radii || !packed_grads || !v_means3d || !v_scales ||
        !v_quats || !v_dc || (K > 1 && !v_rest))
        return TS_ERR_INVALID;
    if (!ts::aligned16(means3d) || !ts::aligned16(scales) || !ts::aligned16(quats) || !ts::aligned16(packed_grads) ||
        !ts::aligned16(v_means3d) || !ts::aligned16(v_scales) || !ts::aligned16(v_quats) ||
        (v_xys_out && (reinterpret_cast<uintptr_t>(v_xys_out) & 7u)))
        return TS_ERR_ALIGN;
    const int grid = (N + ts::kProjThreads - 1) / ts::kProjThreads;
    const size_t smem = sizeof(float) * ts::kProjThreads * ((size_t)(K - 1) * 3 + 3 + 6) + 16;
    cudaStream_t st = (cudaStream_t)stream;
#define TS_LAUNCH_PSB(D)                                                                                        \
    do {                                                                                                        \
        TS_CHECK_CUDA(cudaFuncSetAttribute(ts::project_sh_bwd_kernel<D>,                                        \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),             \
                      "ts_project_sh_bwd/attr");                                                                \
        ts::project_sh_bwd_kernel<D><<<grid, ts::kProjThreads, smem, st>>>(                                     \
            N, K, means3d, scales, glob_scale, (const float4*)quats, viewmat, projmat, fx, fy, cx, cy,          \
            img_height, img_width, flags, radii, (const float4*)packed_grads, opacity_logits, clamp_mask,       \
            v_means3d, v_scales, (float4*)v_quats, v_opacity_logits, (float2*)v_xys_out, v_dc, v_rest);         \
    } while (0)
    switch (degree) {
        case 0: TS_LAUNCH_PSB(0); break;
        case 1: TS_LAUNCH_PSB(1); break;
        case 2: TS_LAUNCH_PSB(2); break;
        case 3: TS_LAUNCH_PSB(3); break;
        default: TS_LAUNCH_PSB(4); break;
    }
#undef TS_LAUNCH_PSB
    TS_CHECK_LAUNCH("ts_project_sh_bwd");
    return TS_OK;
}

static int launch_project_bwd_views(bool compact, int n_views, int N, const float* means3d, const float* scales,
                                    float glob_scale, const float* quats, const float* cams, int img_height,
                                    int img_width, int flags, const float* rows, int64_t view_stride_floats,
                                    const float* opacity_logits, float out_scale, int n_dst, int first_dst,
                                    const ts::PeerPtrs& dm, const ts::PeerPtrs& ds, const ts::PeerPtrs& dq,
                                    const ts::PeerPtrs& dl, ts_stream_t stream) {
    if (n_views < 1 || N < 0 || img_height <= 0 || img_width <= 0 || view_stride_floats < 0 ||
        (view_stride_floats % 4) != 0 || n_dst < 1 || n_dst > ts::kMaxPeers)
        return TS_ERR_INVALID;
    if (N == 0) return TS_OK;
    if (!means3d || !scales || !quats || !cams || !rows) return TS_ERR_INVALID;
    if (!ts::aligned16(means3d) || !ts::aligned16(scales) || !ts::aligned16(quats) || !ts::aligned16(rows))
        return TS_ERR_ALIGN;
    for (int d = 0; d < n_dst; ++d) {
        if (!dm.p[d] || !ds.p[d] || !dq.p[d]) return TS_ERR_INVALID;
        if (!ts::aligned16(dm.p[d]) || !ts::aligned16(ds.p[d]) || !ts::aligned16(dq.p[d])) return TS_ERR_ALIGN;
    }
    int grid = (N + ts::kProjThreads - 1) / ts::kProjThreads;
    cudaStream_t st = (cudaStream_t)stream;
    if (compact)
        ts::project_bwd_views_kernel<true><<<grid, ts::kProjThreads, 0, st>>>(
            n_views, N, means3d, scales, glob_scale, (const float4*)quats, cams, img_height, img_width, flags,
            (const float4*)rows, (size_t)(view_stride_floats / 4), opacity_logits, out_scale, n_dst, first_dst,
            dm, ds, dq, dl);
    else
        ts::project_bwd_views_kernel<false><<<grid, ts::kProjThreads, 0, st>>>(
            n_views, N, means3d, scales, glob_scale, (const float4*)quats, cams, img_height, img_width, flags,
            (const float4*)rows, (size_t)(view_stride_floats / 4), opacity_logits, out_scale, n_dst, first_dst,
            dm, ds, dq, dl);
    TS_CHECK_LAUNCH("ts_project_bwd_views");
    return TS_OK;
}

int ts_project_bwd_views(int n_views, int N, const float* means3d, const float* scales, float glob_scale,
                         const float* quats, const float* cams, int img_height, int img_width, int flags,
                         const float* packed_grads, int64_t view_stride_floats, const float* opacity_logits,
                         float out_scale, float* v_means3d, float* v_scales, float* v_quats,
                         float* v_opacity_logits, ts_stream_t stream) {
    ts::PeerPtrs dm{}, ds{}, dq{}, dl{};
    dm.p[0] = v_means3d; ds.p[0] = v_scales; dq.p[0] = v_quats; dl.p[0] = v_opacity_logits;
    return launch_project_bwd_views(false, n_views, N, means3d, scales, glob_scale, quats, cams, img_height,
                                    img_width, flags, packed_grads, view_stride_floats, opacity_logits, out_scale,
                                    1, 0, dm, ds, dq, dl, stream);
}

int ts_project_bwd_views_peer(int n_views, int N, const float* means3d, const float* scales, float glob_scale,
                              const float* quats, const float* cams, int img_height, int img_width, int flags,
                              const float* geo_rows, int64_t view_stride_floats, const float* opacity_logits,
                              float out_scale, int n_dst, int first_dst, void* const* v_means_ptrs_host,
                              void* const* v_scales_ptrs_host, void* const* v_quats_ptrs_host,
                              void* const* v_logit_ptrs_host, ts_stream_t stream) {
    if (n_dst < 1 || n_dst > ts::kMaxPeers || !v_means_ptrs_host || !v_scales_ptrs_host || !v_quats_ptrs_host)
        return TS_ERR_INVALID;
    ts::PeerPtrs dm{}, ds{}, dq{}, dl{};
    for (int d = 0; d < n_dst; ++d) {
        dm.p[d] = v_means_ptrs_host[d]; ds.p[d] = v_scales_ptrs_host[d]; dq.p[d] = v_quats_ptrs_host[d];
        dl.p[d] = v_logit_ptrs_host ? v_logit_ptrs_host[d] : nullptr;
    }
    return launch_project_bwd_views(true, n_views, N, means3d, scales, glob_scale, quats, cams, img_height,
                                    img_width, flags, geo_rows, view_stride_floats, opacity_logits, out_scale,
                                    n_dst, first_dst, dm, ds, dq, dl, stream);
}

}  // extern "C"
#endif  // !TS_HOST_EMU
