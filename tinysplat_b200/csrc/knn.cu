// SURVEY.md 8(f)-3 — K nearest neighbours for the density regularizer.
// Stand-in for pytorch3d.ops.knn_points as the reference calls it:
// knn_points(points[None], means[None], K=16).idx[0]  [REF tinysplat/splatting/model_gaussian.py:260,425,519].
// Brute force, exact: one thread per query keeps its K best (squared distance, index) pairs sorted
// in registers; reference points stream through shared memory in tiles and are read by broadcast.
// Runs only when the neighbour lists are refreshed (every densify interval), not every step.
#include "ts_common.cuh"

namespace ts {

constexpr int kKnnThreads = 128;
constexpr int kKnnTile = 1024;

template <int K>
__global__ void __launch_bounds__(kKnnThreads)
knn_kernel(int P1, int P2, const float* __restrict__ q, const float* __restrict__ ref,
           float* __restrict__ dists, int64_t* __restrict__ idx) {
    __shared__ float4 s_ref[kKnnTile];
    const int i = blockIdx.x * kKnnThreads + threadIdx.x;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (i < P1) { qx = __ldg(q + 3 * (size_t)i); qy = __ldg(q + 3 * (size_t)i + 1); qz = __ldg(q + 3 * (size_t)i + 2); }
    float bd[K];
    int bi[K];
#pragma unroll
    for (int t = 0; t < K; ++t) { bd[t] = __int_as_float(0x7f800000); bi[t] = -1; }
    for (int t0 = 0; t0 < P2; t0 += kKnnTile) {
        const int nt = min(kKnnTile, P2 - t0);
        for (int j = threadIdx.x; j < nt; j += kKnnThreads) {
            const float* r = ref + 3 * (size_t)(t0 + j);
            s_ref[j] = make_float4(__ldg(r), __ldg(r + 1), __ldg(r + 2), 0.f);
        }
        __syncthreads();
#pragma unroll 4
        for (int j = 0; j < nt; ++j) {
            const float4 r = s_ref[j];
            float dx = qx - r.x, dy = qy - r.y, dz = qz - r.z;
            float d = dx * dx + dy * dy + dz * dz;
            if (d < bd[K - 1]) {            // rare after warm-up: ~K ln(P2/K) insertions per query
                bd[K - 1] = d; bi[K - 1] = t0 + j;
#pragma unroll
                for (int t = K - 1; t > 0; --t) {
                    if (bd[t] < bd[t - 1]) {
                        float td = bd[t]; bd[t] = bd[t - 1]; bd[t - 1] = td;
                        int ti = bi[t]; bi[t] = bi[t - 1]; bi[t - 1] = ti;
                    }
                }
            }
        }
        __syncthreads();
    }
    if (i < P1) {
#pragma unroll
        for (int t = 0; t < K; ++t) {
            dists[(size_t)i * K + t] = (bi[t] >= 0) ? bd[t] : 0.f;
            idx[(size_t)i * K + t] = (bi[t] >= 0) ? (int64_t)bi[t] : 0;   // fewer than K references: pad with 0
        }
    }
}

}  // namespace ts

#ifndef TS_HOST_EMU
extern "C" {

int ts_knn_points(int P1, int P2, int K, const float* queries, const float* refs, float* dists, int64_t* idx,
                  ts_stream_t stream) {
    if (P1 < 0 || P2 < 0 || K < 1) return TS_ERR_INVALID;
    if (P1 == 0) return TS_OK;
    if (!queries || !dists || !idx || (P2 > 0 && !refs)) return TS_ERR_INVALID;
    int grid = (P1 + ts::kKnnThreads - 1) / ts::kKnnThreads;
    cudaStream_t st = (cudaStream_t)stream;
#define TS_LAUNCH_KNN(KK) ts::knn_kernel<KK><<<grid, ts::kKnnThreads, 0, st>>>(P1, P2, queries, refs, dists, idx)
    switch (K) {
        case 1: TS_LAUNCH_KNN(1); break;
        case 2: TS_LAUNCH_KNN(2); break;
        case 4: TS_LAUNCH_KNN(4); break;
        case 8: TS_LAUNCH_KNN(8); break;
        case 16: TS_LAUNCH_KNN(16); break;
        case 32: TS_LAUNCH_KNN(32); break;
        default: return TS_ERR_INVALID;     // K in {1, 2, 4, 8, 16, 32}
    }
#undef TS_LAUNCH_KNN
    TS_CHECK_LAUNCH("ts_knn_points");
    return TS_OK;
}

}  // extern "C"
#endif  // !TS_HOST_EMU
