// Shared device helpers and constants of the tinysplat_b200 kernels (sm_100a only).
#pragma once
// TS_HOST_EMU: tests/emu compiles the grouped blend kernels as HOST code on a fiber-based SIMT
// emulator (threadIdx, __shared__, warp votes/shuffles, barriers) so that the kernel logic is
// checked against the oracle without a GPU.  Test infrastructure only; never defined in the product.
#ifdef TS_HOST_EMU
#include "ts_emu.h"
#else
#include <cuda_runtime.h>
#endif
#include <stdint.h>
#include "../../include/tinysplat_b200.h"

// Dynamic shared memory of a kernel.  The emulator has no launch-time size: it gets a static
// buffer of the largest size any kernel asks for.
#ifdef TS_HOST_EMU
#define TS_DYN_SMEM(type, name, align) \
    static __attribute__((aligned(align))) type name[(160 * 1024) / sizeof(type)]
#else
#define TS_DYN_SMEM(type, name, align) extern __shared__ __align__(align) type name[]
#endif

namespace ts {

// ---- constants of the stated algorithm (mirrored in oracle/gsplat_oracle.py) --------------
constexpr int   kBlock      = 16;        // tile edge [REF rasterize.py:19-20]
constexpr float kCov2dBlur  = 0.3f;
constexpr float kEigFloor   = 0.1f;
constexpr float kFovClamp   = 1.3f;
constexpr float kAlphaMax   = 0.999f;
constexpr float kAlphaMin   = 1.0f / 255.0f;
constexpr float kTStop      = 1e-4f;
constexpr float kWEps       = 1e-6f;
constexpr float kPixCenter  = 0.5f;
constexpr float kLog2e      = 1.4426950408889634f;
constexpr float kMaxRadius  = 1073741824.0f;  // 2^30

// Packed raster record: 3 x float4 per Gaussian, 48 B, gathered by the blend kernels.
//   q0 = {x, y, hx, hy}   centre (pixels) and half-extents of the alpha>=1/255 footprint
//   q1 = {A, B, C, opac}  conic pre-scaled so that opac*exp2(-(A dx^2 + B dx dy + C dy^2))
//                         equals opac*exp(-sigma):  A = .5*log2e*a, B = log2e*b, C = .5*log2e*c
//   q2 = {c0, c1, c2, c3} colour channels (unused ones are zero)
constexpr int kRecFloats  = 12;
// Packed gradient record accumulated by blend-backward (3 x float4):
//   g0 = {S_x, S_y, S_xx, S_xy}  sums of v_sigma*dx, v_sigma*dy, v_sigma*dx^2, v_sigma*dx*dy
//   g1 = {S_yy, v_opac, 0, 0}
//   g2 = {v_c0, v_c1, v_c2, v_c3}
constexpr int kGradFloats = 12;

void set_last_error(const char* where, cudaError_t e);
void count_launch(int n = 1);

#define TS_CHECK_LAUNCH(where)                                     \
    do {                                                           \
        cudaError_t e__ = cudaGetLastError();                      \
        if (e__ != cudaSuccess) {                                  \
            ts::set_last_error(where, e__);                        \
            return TS_ERR_CUDA;                                    \
        }                                                          \
        ts::count_launch();                                        \
    } while (0)

#define TS_CHECK_CUDA(expr, where)                                 \
    do {                                                           \
        cudaError_t e__ = (expr);                                  \
        if (e__ != cudaSuccess) {                                  \
            ts::set_last_error(where, e__);                        \
            return TS_ERR_CUDA;                                    \
        }                                                          \
    } while (0)

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
__device__ __forceinline__ bool aligned_dev16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- block-cooperative, 128-bit vectorised global <-> shared staging ---------------------
// A block of THREADS threads owns items [item0, item0+THREADS) of an [n_items, K] fp32 array.
// item0 is a multiple of THREADS, so item0*K*4 bytes keeps the 16 B alignment of the base.
template <int K, int THREADS>
__device__ __forceinline__ void block_load(const float* __restrict__ g, float* s, int item0,
                                           int n_items) {
    int n_valid = min(THREADS, n_items - item0);
    int nfl = n_valid * K;
    const float* src = g + (size_t)item0 * K;
    int nv4 = nfl >> 2;
    const float4* src4 = reinterpret_cast<const float4*>(src);
    float4* s4 = reinterpret_cast<float4*>(s);
    for (int i = threadIdx.x; i < nv4; i += THREADS) s4[i] = __ldg(src4 + i);
    for (int i = (nv4 << 2) + threadIdx.x; i < nfl; i += THREADS) s[i] = __ldg(src + i);
}

template <int K, int THREADS>
__device__ __forceinline__ void block_store(float* __restrict__ g, const float* s, int item0,
                                            int n_items) {
    int n_valid = min(THREADS, n_items - item0);
    int nfl = n_valid * K;
    float* dst = g + (size_t)item0 * K;
    int nv4 = nfl >> 2;
    float4* dst4 = reinterpret_cast<float4*>(dst);
    const float4* s4 = reinterpret_cast<const float4*>(s);
    for (int i = threadIdx.x; i < nv4; i += THREADS) dst4[i] = s4[i];
    for (int i = (nv4 << 2) + threadIdx.x; i < nfl; i += THREADS) dst[i] = s[i];
}

// (w,x,y,z) quaternion -> row-major rotation matrix, used as given (caller normalises).
__device__ __forceinline__ void quat_to_rotmat(float4 q, float R[9]) {
    float w = q.x, x = q.y, y = q.z, z = q.w;
    R[0] = 1.f - 2.f * (y * y + z * z); R[1] = 2.f * (x * y - w * z); R[2] = 2.f * (x * z + w * y);
    R[3] = 2.f * (x * y + w * z); R[4] = 1.f - 2.f * (x * x + z * z); R[5] = 2.f * (y * z - w * x);
    R[6] = 2.f * (x * z - w * y); R[7] = 2.f * (y * z + w * x); R[8] = 1.f - 2.f * (x * x + y * y);
}

// Tile rectangle [lo, hi) of the square of half-edge `radius` around (x, y), clipped to the
// tile grid.  fp32 arithmetic mirrors oracle tile_bbox().
__device__ __forceinline__ void tile_bbox(float x, float y, float radius, int tbx, int tby,
                                          int& lox, int& loy, int& hix, int& hiy) {
    float cx = x / (float)kBlock, cy = y / (float)kBlock, r = radius / (float)kBlock;
    float fx0 = floorf(fminf(fmaxf(cx - r, -1e9f), 1e9f));
    float fy0 = floorf(fminf(fmaxf(cy - r, -1e9f), 1e9f));
    float fx1 = floorf(fminf(fmaxf(cx + r + 1.f, -1e9f), 1e9f));
    float fy1 = floorf(fminf(fmaxf(cy + r + 1.f, -1e9f), 1e9f));
    lox = min(max(0, (int)fx0), tbx); loy = min(max(0, (int)fy0), tby);
    hix = min(max(0, (int)fx1), tbx); hiy = min(max(0, (int)fy1), tby);
}

__device__ __forceinline__ float warp_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 16);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}

#ifdef TS_HOST_EMU
__device__ __forceinline__ float ex2_approx(float x) { return exp2f(x); }
__device__ __forceinline__ float rcp_approx(float x) { return 1.f / x; }
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) { memcpy(smem_dst, gmem_src, 16); }
__device__ __forceinline__ void cp_async_commit() {}
template <int N>
__device__ __forceinline__ void cp_async_wait() {}
// TMA bulk copies: executed synchronously by the issuing thread, so every wait is trivially
// satisfied once the block barrier that follows the issue has been passed
__device__ __forceinline__ void mbar_init(uint64_t*, unsigned) {}
__device__ __forceinline__ void mbar_expect_tx(uint64_t*, unsigned) {}
__device__ __forceinline__ void mbar_wait(uint64_t*, unsigned) {}
__device__ __forceinline__ void fence_proxy_async() {}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, uint64_t*) { memcpy(smem_dst, gmem_src, bytes); }
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, unsigned bytes) { memcpy(gmem_dst, smem_src, bytes); }
__device__ __forceinline__ void bulk_commit() {}
__device__ __forceinline__ void bulk_wait_read0() {}
__device__ __forceinline__ void bulk_wait_read1() {}
__device__ __forceinline__ void bulk_wait0() {}
#else
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ float rcp_approx(float x) {   // MUFU.RCP, <= 1 ulp
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// 16-byte async global->shared copy (LDGSTS), commit / wait.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    uint32_t s = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

// ---- TMA bulk copies (cp.async.bulk, SASS: UBLKCP) + mbarrier ------------------------------
// One elected thread moves a whole contiguous block between global and shared memory; the
// copy engine, not the SM's LSU, generates the addresses.  Sizes/addresses: multiples of 16 B.
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, uint64_t* bar) {
    uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    uint32_t b = (uint32_t)__cvta_generic_to_shared(bar);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(d), "l"(gmem_src), "r"(bytes), "r"(b) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, unsigned bytes) {
    uint32_t s = (uint32_t)__cvta_generic_to_shared(smem_src);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(s), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// at most one bulk group of this thread may still be reading its shared-memory source
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
// all bulk groups of this thread have COMPLETED (writes performed), not merely read their source
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

#endif  // TS_HOST_EMU

}  // namespace ts
