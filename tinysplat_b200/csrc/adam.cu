// SURVEY.md 8(f)-2 — fused multi-tensor Adam step for the Gaussian parameters.
// The reference trains with torch.optim.Adam over six parameter groups
// [REF scripts/train.py:26; tinysplat/splatting/model_gaussian.py:112-120] — 59 floats per
// Gaussian, the next-largest per-step HBM consumer after the rasterizer.  One launch updates all
// tensors: 16 B read + 12 B written per element (param, grad, exp_avg, exp_avg_sq), 128-bit
// accesses, grid-stride over a flat (tensor, chunk) work list.  Arithmetic follows
// torch.optim.Adam (no weight decay, no amsgrad) operation for operation.
#include <cmath>
#include "ts_common.cuh"

namespace ts {

constexpr int kAdamMaxTensors = 8;
constexpr int kAdamThreads = 256;
constexpr int kAdamVecPerThread = 4;                       // float4 per thread per block-iteration
constexpr int kAdamChunk = kAdamThreads * kAdamVecPerThread * 4;   // floats per block

struct AdamTensors {
    float* param[kAdamMaxTensors];
    const float* grad[kAdamMaxTensors];
    float* m[kAdamMaxTensors];
    float* v[kAdamMaxTensors];
    long long n[kAdamMaxTensors];
    float step_size[kAdamMaxTensors];      // lr / (1 - beta1^t)
    float sqrt_bc2[kAdamMaxTensors];       // sqrt(1 - beta2^t)
    int first_block[kAdamMaxTensors + 1];  // prefix of per-tensor block counts
    int count;
};

// omb1 / omb2 are (1 - beta) formed in DOUBLE on the host, as torch does: 1.f - 0.999f is off by
// 1.3e-5 relative in fp32.
__device__ __forceinline__ void adam_update(float& p, float g, float& m, float& v, float omb1, float b2,
                                            float omb2, float eps, float step_size, float sqrt_bc2) {
    m = m + omb1 * (g - m);                            // exp_avg.lerp_(grad, 1 - beta1)
    v = v * b2 + omb2 * (g * g);                       // exp_avg_sq.mul_(beta2).addcmul_(g, g, 1 - beta2)
    float denom = sqrtf(v) / sqrt_bc2 + eps;           // (sqrt(v) / sqrt(bc2)) + eps
    p = p - step_size * (m / denom);                   // param.addcdiv_(exp_avg, denom, -step_size)
}

__global__ void __launch_bounds__(kAdamThreads)
adam_multi_kernel(AdamTensors t, float omb1, float b2, float omb2, float eps) {
    // which tensor does this block belong to?
    int ti = 0;
#pragma unroll
    for (int k = 1; k < kAdamMaxTensors; ++k)
        if (k < t.count && (int)blockIdx.x >= t.first_block[k]) ti = k;
    const long long base = (long long)((int)blockIdx.x - t.first_block[ti]) * kAdamChunk;
    const long long n = t.n[ti];
    float* __restrict__ P = t.param[ti];
    const float* __restrict__ G = t.grad[ti];
    float* __restrict__ M = t.m[ti];
    float* __restrict__ V = t.v[ti];
    const float ss = t.step_size[ti], ib = t.sqrt_bc2[ti];
    const bool vec_ok = (((uintptr_t)P | (uintptr_t)G | (uintptr_t)M | (uintptr_t)V) & 15u) == 0;
#pragma unroll
    for (int it = 0; it < kAdamVecPerThread; ++it) {
        long long i = base + ((long long)it * kAdamThreads + threadIdx.x) * 4;
        if (i >= n) break;
        if (vec_ok && i + 4 <= n) {
            float4 p = *reinterpret_cast<float4*>(P + i);
            float4 g = __ldg(reinterpret_cast<const float4*>(G + i));
            float4 m = *reinterpret_cast<float4*>(M + i);
            float4 v = *reinterpret_cast<float4*>(V + i);
            adam_update(p.x, g.x, m.x, v.x, omb1, b2, omb2, eps, ss, ib);
            adam_update(p.y, g.y, m.y, v.y, omb1, b2, omb2, eps, ss, ib);
            adam_update(p.z, g.z, m.z, v.z, omb1, b2, omb2, eps, ss, ib);
            adam_update(p.w, g.w, m.w, v.w, omb1, b2, omb2, eps, ss, ib);
            *reinterpret_cast<float4*>(P + i) = p;
            *reinterpret_cast<float4*>(M + i) = m;
            *reinterpret_cast<float4*>(V + i) = v;
        } else {
            for (long long j = i; j < n && j < i + 4; ++j) {
                float p = P[j], m = M[j], v = V[j];
                adam_update(p, G[j], m, v, omb1, b2, omb2, eps, ss, ib);
                P[j] = p; M[j] = m; V[j] = v;
            }
        }
    }
}

// Host side of ts_adam_step: validates the arguments and builds the flat (tensor, block) work list.
// Returns TS_OK with blocks == 0 when there is nothing to do.
static int adam_build(int num_tensors, float* const* params, const float* const* grads, float* const* exp_avgs,
                      float* const* exp_avg_sqs, const int64_t* numels, const float* lrs, const int64_t* steps,
                      double beta1, double beta2, AdamTensors& t, int& blocks) {
    blocks = 0;
    t.count = 0;
    if (num_tensors < 0 || num_tensors > kAdamMaxTensors) return TS_ERR_INVALID;
    if (num_tensors == 0) return TS_OK;
    if (!params || !grads || !exp_avgs || !exp_avg_sqs || !numels || !lrs || !steps) return TS_ERR_INVALID;
    for (int k = 0; k < num_tensors; ++k) {
        if (numels[k] < 0 || steps[k] < 1) return TS_ERR_INVALID;
        if (numels[k] == 0) continue;
        if (!params[k] || !grads[k] || !exp_avgs[k] || !exp_avg_sqs[k]) return TS_ERR_INVALID;
        int c = t.count++;
        t.param[c] = params[k]; t.grad[c] = grads[k]; t.m[c] = exp_avgs[k]; t.v[c] = exp_avg_sqs[k];
        t.n[c] = numels[k];
        double bc1 = 1.0 - pow(beta1, (double)steps[k]);
        double bc2 = 1.0 - pow(beta2, (double)steps[k]);
        t.step_size[c] = (float)((double)lrs[k] / bc1);
        t.sqrt_bc2[c] = (float)sqrt(bc2);
        t.first_block[c] = blocks;
        blocks += (int)((numels[k] + kAdamChunk - 1) / kAdamChunk);
    }
    t.first_block[t.count] = blocks;
    for (int c = t.count + 1; c <= kAdamMaxTensors; ++c) t.first_block[c] = blocks;
    return TS_OK;
}

}  // namespace ts

#ifndef TS_HOST_EMU
extern "C" {

int ts_adam_max_tensors(void) { return ts::kAdamMaxTensors; }

int ts_adam_step(int num_tensors, float* const* params, const float* const* grads, float* const* exp_avgs,
                 float* const* exp_avg_sqs, const int64_t* numels, const float* lrs, const int64_t* steps,
                 double beta1, double beta2, double eps, ts_stream_t stream) {
    ts::AdamTensors t;
    int blocks = 0;
    int rc = ts::adam_build(num_tensors, params, grads, exp_avgs, exp_avg_sqs, numels, lrs, steps, beta1, beta2, t,
                            blocks);
    if (rc != TS_OK || blocks == 0) return rc;
    ts::adam_multi_kernel<<<blocks, ts::kAdamThreads, 0, (cudaStream_t)stream>>>(
        t, (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2), (float)eps);
    TS_CHECK_LAUNCH("ts_adam_step");
    return TS_OK;
}

}  // extern "C"
#endif  // !TS_HOST_EMU
