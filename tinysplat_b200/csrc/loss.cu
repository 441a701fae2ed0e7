// Fused L1 image loss for the training step around the rasterizer (the caller side of
// the hot path): loss = mean |rendered - ground truth|  [REF scripts/train.py:58-59], forward and the
// gradient in ONE pass over the image.  The ground truth may be the uint8 image as it is stored on the
// host (the reference keeps `uint8 / 255` as a float32 CPU tensor and uploads 4 bytes per channel
// every step [REF tinysplat/scene.py:27-31,130-132]): value = u8 / 255 formed here, so a step uploads
// a quarter of the bytes.
// HBM-bound streaming kernel: reads 4 B (image) + 1 or 4 B (target), writes 4 B (gradient) per
// element; block partial sums in a fixed order -> the loss is deterministic (no float atomics).
#include "ts_common.cuh"

namespace ts {

constexpr int kLossThreads = 256;
constexpr int kLossMaxBlocks = 1184;     // 148 SMs x 8 resident CTAs: one wave, grid-stride over the image

template <typename T>
__device__ __forceinline__ float4 load_target4(const T* __restrict__ t, size_t i4);
template <>
__device__ __forceinline__ float4 load_target4<float>(const float* __restrict__ t, size_t i4) {
    return __ldg(reinterpret_cast<const float4*>(t) + i4);
}
template <>
__device__ __forceinline__ float4 load_target4<uint8_t>(const uint8_t* __restrict__ t, size_t i4) {
    const uchar4 u = __ldg(reinterpret_cast<const uchar4*>(t) + i4);
    // the same value torch computes for `uint8_tensor / 255` (a correctly rounded fp32 division)
    return make_float4(__fdiv_rn((float)u.x, 255.f), __fdiv_rn((float)u.y, 255.f), __fdiv_rn((float)u.z, 255.f),
                       __fdiv_rn((float)u.w, 255.f));
}
template <typename T>
__device__ __forceinline__ float load_target1(const T* __restrict__ t, size_t i);
template <>
__device__ __forceinline__ float load_target1<float>(const float* __restrict__ t, size_t i) { return __ldg(t + i); }
template <>
__device__ __forceinline__ float load_target1<uint8_t>(const uint8_t* __restrict__ t, size_t i) {
    return __fdiv_rn((float)__ldg(t + i), 255.f);
}

__device__ __forceinline__ float sgn(float d) { return d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f); }

// grad[i] = sign(img[i] - target[i]) * grad_scale  (grad_scale = 1 / n for the mean);
// partials[block] = sum over the block's elements of |img - target|; the LAST block to finish adds the
// partials in index order and writes loss[0] = sum * loss_scale.  counter must be 0 on entry and is
// reset to 0 on exit.
template <typename T>
__global__ void __launch_bounds__(kLossThreads)
l1_loss_kernel(size_t n, const float* __restrict__ img, const T* __restrict__ target, float grad_scale,
               float loss_scale, float* __restrict__ grad, float* __restrict__ partials,
               unsigned int* __restrict__ counter, float* __restrict__ loss) {
    __shared__ float s_warp[kLossThreads / 32];
    __shared__ bool s_last;
    const size_t n4 = n >> 2;
    float acc = 0.f;
    for (size_t i4 = (size_t)blockIdx.x * kLossThreads + threadIdx.x; i4 < n4; i4 += (size_t)gridDim.x * kLossThreads) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(img) + i4);
        const float4 t = load_target4<T>(target, i4);
        const float dx = a.x - t.x, dy = a.y - t.y, dz = a.z - t.z, dw = a.w - t.w;
        acc += (fabsf(dx) + fabsf(dy)) + (fabsf(dz) + fabsf(dw));
        if (grad)
            reinterpret_cast<float4*>(grad)[i4] = make_float4(sgn(dx) * grad_scale, sgn(dy) * grad_scale,
                                                              sgn(dz) * grad_scale, sgn(dw) * grad_scale);
    }
    if (blockIdx.x == 0) {      // the up-to-3-element tail
        for (size_t i = (n4 << 2) + threadIdx.x; i < n; i += kLossThreads) {
            const float d = __ldg(img + i) - load_target1<T>(target, i);
            acc += fabsf(d);
            if (grad) grad[i] = sgn(d) * grad_scale;
        }
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float b = 0.f;
#pragma unroll
        for (int w = 0; w < kLossThreads / 32; ++w) b += s_warp[w];
        partials[blockIdx.x] = b;
        __threadfence();
        s_last = atomicAdd(counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // last block: fixed-order sum of the partials (a thread sums a strided subset, then the same tree)
    float tot = 0.f;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += kLossThreads) tot += partials[b];
    tot = warp_sum(tot);
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = tot;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kLossThreads / 32; ++w) s += s_warp[w];
        loss[0] = s * loss_scale;
        *counter = 0u;
    }
}

}  // namespace ts

#ifndef TS_HOST_EMU
extern "C" {

int ts_l1_loss_work_floats(void) { return ts::kLossMaxBlocks + 4; }

int ts_l1_loss(int64_t n, const float* img, const void* target, int target_is_u8, float grad_scale,
               float loss_scale, float* grad, float* work, float* loss, ts_stream_t stream) {
    if (n <= 0 || !img || !target || !work || !loss) return TS_ERR_INVALID;
    if (!ts::aligned16(img) || (grad && !ts::aligned16(grad)) ||
        (reinterpret_cast<uintptr_t>(target) & (target_is_u8 ? 3u : 15u)))
        return TS_ERR_ALIGN;
    const int64_t n4 = n >> 2;
    int grid = (int)((n4 + ts::kLossThreads - 1) / ts::kLossThreads);
    grid = grid < 1 ? 1 : (grid > ts::kLossMaxBlocks ? ts::kLossMaxBlocks : grid);
    unsigned int* counter = reinterpret_cast<unsigned int*>(work + ts::kLossMaxBlocks);   // zeroed once by the caller
    cudaStream_t st = (cudaStream_t)stream;
    if (target_is_u8)
        ts::l1_loss_kernel<uint8_t><<<grid, ts::kLossThreads, 0, st>>>((size_t)n, img, (const uint8_t*)target, grad_scale,
                                                                      loss_scale, grad, work, counter, loss);
    else
        ts::l1_loss_kernel<float><<<grid, ts::kLossThreads, 0, st>>>((size_t)n, img, (const float*)target, grad_scale,
                                                                    loss_scale, grad, work, counter, loss);
    TS_CHECK_LAUNCH("ts_l1_loss");
    return TS_OK;
}

}  // extern "C"
#endif  // !TS_HOST_EMU
