// K3 — tile binning and per-tile depth sort.
// Happens behind gsplat.rasterize_gaussians  [REF tinysplat/splatting/rasterize.py:44,50].
//
// B200-first design (instead of one global 64-bit radix sort over all intersections):
//   count  : per Gaussian, atomically count the tiles its alpha>=1/255 footprint can reach
//            (3-sigma bbox of the stated algorithm, intersected with the opacity-aware extent:
//            a pair that cannot light any pixel is never emitted -> smaller M, same image);
//            the same pass packs the 48-byte raster record the blend kernels gather.
//   scan   : exclusive scan of per-tile counts (single CTA; T <= ~130k tiles).
//   emit   : scatter key = depth_bits<<32 | gaussian_id into the tile's bucket.
//   sort   : per tile, on 64-bit keys (unique -> deterministic, ties in depth resolved by gaussian id
//            exactly like a stable sort of the emission order): lists of <= 512 entries by ONE WARP
//            in registers; longer lists by one CTA — every warp sorts a chunk in registers, then
//            log2(#warps) merge-path levels in shared memory (O(n log n) instead of the O(n log^2 n)
//            of a shared-memory bitonic network); writes the 32-bit id list.
// Traffic per intersection: 8 B write + 8 B read + 4 B write, vs ~150 B for 6 radix passes.
#include <climits>
#include "ts_common.cuh"
#include "ts_binning.cuh"

namespace ts {

constexpr int kBinThreads = 256;
constexpr int kSmemSortCap = 16384;  // 128 KB of 64-bit keys
template <int CH>
__global__ void __launch_bounds__(kBinThreads)
bin_count_kernel(int N, const float2* __restrict__ xys, const int32_t* __restrict__ radii,
                 const float* __restrict__ conics, const float* __restrict__ opacity,
                 const float* __restrict__ colors, int tbx, int tby, int cull, int flags,
                 float4* __restrict__ recs, int32_t* __restrict__ tile_counts) {
    const int i = blockIdx.x * kBinThreads + threadIdx.x;
    int lox = 0, loy = 0, hix = 0, hiy = 0;
    if (i < N) {
        int r = __ldg(radii + i);
        float2 xy = __ldg(xys + i);
        float a = __ldg(conics + 3 * i), b = __ldg(conics + 3 * i + 1), c = __ldg(conics + 3 * i + 2);
        float op = __ldg(opacity + i);
        if (flags & TS_BIN_OPACITY_LOGIT) op = 1.f / (1.f + expf(-op));   // sigmoid [REF rasterize.py:86]
        float hx, hy;
        footprint_extent(a, b, c, op, cull, hx, hy);
        float4 q0 = make_float4(xy.x, xy.y, hx, hy);
        float4 q1 = make_float4(0.5f * kLog2e * a, kLog2e * b, 0.5f * kLog2e * c, op);
        recs[3 * (size_t)i] = q0;
        recs[3 * (size_t)i + 1] = q1;
        if (colors) {   // otherwise the colour float4 is written by ts_sh_fwd (fused pipeline)
            float4 q2 = make_float4(0.f, 0.f, 0.f, 0.f);
            q2.x = __ldg(colors + (size_t)CH * i);
            if (CH > 1) q2.y = __ldg(colors + (size_t)CH * i + 1);
            if (CH > 2) q2.z = __ldg(colors + (size_t)CH * i + 2);
            if (CH > 3) q2.w = __ldg(colors + (size_t)CH * i + 3);
            recs[3 * (size_t)i + 2] = q2;
        }
        if (r > 0 && !(flags & TS_BIN_PACK_ONLY)) tile_rect(q0, (float)r, tbx, tby, cull, lox, loy, hix, hiy);
    }
    if (flags & TS_BIN_PACK_ONLY) return;       // uniform: the tile lists of an earlier call are reused
    for_each_tile(lox, loy, hix, hiy, tbx, 0u, 0u,
                  [&](int tile, uint32_t, uint32_t) { atomicAdd(tile_counts + (size_t)tile * kCounterStride, 1); });
}

__global__ void __launch_bounds__(kBinThreads)
bin_emit_kernel(int N, const float* __restrict__ depths, const int32_t* __restrict__ radii,
                const float4* __restrict__ recs, int tbx, int tby, int cull,
                int32_t* __restrict__ cursors, uint64_t* __restrict__ keys, int cap) {
    const int i = blockIdx.x * kBinThreads + threadIdx.x;
    int lox = 0, loy = 0, hix = 0, hiy = 0;
    uint64_t key = 0;
    if (i < N) {
        int r = __ldg(radii + i);
        if (r > 0) {
            float4 q0 = __ldg(recs + 3 * (size_t)i);
            tile_rect(q0, (float)r, tbx, tby, cull, lox, loy, hix, hiy);
            key = ((uint64_t)__float_as_uint(__ldg(depths + i)) << 32) | (uint32_t)i;
        }
    }
    // Small footprints (the common case): issue ALL of this Gaussian's return-atomics first,
    // then the dependent key stores, so the ~600-cycle L2 round trips overlap instead of
    // serialising (the kernel was long-scoreboard bound at 12% issue utilisation).
    {
        const int w = hix - lox, h = hiy - loy;
        const int n = (w > 0 && h > 0) ? w * h : 0;
        if (n > 0 && n <= kCoopThreshold) {
            int slots[kCoopThreshold];
            int x = lox, y = loy;
#pragma unroll
            for (int k = 0; k < kCoopThreshold; ++k) {
                if (k < n) {
                    slots[k] = atomicAdd(cursors + (size_t)(y * tbx + x) * kCounterStride, 1);
                    if (++x == hix) { x = lox; ++y; }
                }
            }
#pragma unroll
            for (int k = 0; k < kCoopThreshold; ++k)
                if (k < n && slots[k] < cap) keys[slots[k]] = key;     // cap: see ts_bin_emit
            lox = loy = hix = hiy = 0;   // done; nothing left for the cooperative path
        }
    }
    for_each_tile(lox, loy, hix, hiy, tbx, (uint32_t)key, (uint32_t)(key >> 32),
                  [&](int tile, uint32_t lo, uint32_t hi) {
                      int slot = atomicAdd(cursors + (size_t)tile * kCounterStride, 1);
                      if (slot < cap) keys[slot] = ((uint64_t)hi << 32) | lo;
                  });
}

// Exclusive scan over the tile counts + max / oversize statistics.  Up to kScanMaxBlocks CTAs of
// 1024 threads; each thread owns `chunk` consecutive tiles (all of its strided counter loads in
// flight at once).  A block publishes its aggregate, then sums its predecessors' aggregates in
// parallel (one thread per predecessor, look-back on a flag): no serial chain, no second launch.
// Blocks are dispatched in index order and only ever wait on LOWER indices, so the spin cannot
// deadlock even when other kernels share the GPU.
constexpr int kScanMaxChunk = 16;
constexpr int kScanMaxBlocks = 128;
constexpr int kScanWorkInts = 4 + 2 * kScanMaxBlocks;   // stats[4] | agg[128] | flag[128]

__global__ void __launch_bounds__(1024)
bin_scan_kernel(int T, int chunk, int32_t* __restrict__ counts, int32_t* __restrict__ offsets,
                int32_t* __restrict__ work, int cap) {
    __shared__ int s_warp[32];
    __shared__ int s_total, s_prefix;
    int32_t* stats = work;
    volatile int32_t* agg = work + 4;
    volatile int32_t* flag = work + 4 + kScanMaxBlocks;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int first = (blockIdx.x * 1024 + tid) * chunk;
    int v[kScanMaxChunk];
    int sum = 0, lmax = 0, lbig = 0;
#pragma unroll
    for (int k = 0; k < kScanMaxChunk; ++k) {
        int idx = first + k;
        v[k] = (k < chunk && idx < T) ? counts[(size_t)idx * kCounterStride] : 0;
    }
#pragma unroll
    for (int k = 0; k < kScanMaxChunk; ++k) {
        sum += v[k];
        lmax = max(lmax, v[k]);
        lbig += (v[k] > cap) ? 1 : 0;
    }
    // block-wide inclusive scan of the per-thread sums
    int x = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= d) x += y;
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    if (warp == 0) {
        int wsum = s_warp[lane];
        int xs = wsum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int y = __shfl_up_sync(0xffffffffu, xs, d);
            if (lane >= d) xs += y;
        }
        s_warp[lane] = xs - wsum;
        if (lane == 31) s_total = xs;
    }
    __syncthreads();
    const int excl_in_block = s_warp[warp] + x - sum;
    if (tid == 0) {
        agg[blockIdx.x] = s_total;
        __threadfence();
        flag[blockIdx.x] = 1;
    }
    // look-back: thread t < blockIdx.x fetches predecessor t's aggregate
    int prev = 0;
    if (tid < (int)blockIdx.x) {
        while (flag[tid] == 0) { }
        __threadfence();
        prev = agg[tid];
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) prev += __shfl_xor_sync(0xffffffffu, prev, d);
    __syncthreads();               // s_warp reads above are done
    if (lane == 0) s_warp[warp] = prev;
    __syncthreads();
    if (warp == 0) {
        int p = s_warp[lane];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) p += __shfl_xor_sync(0xffffffffu, p, d);
        if (lane == 0) s_prefix = p;
    }
    __syncthreads();
    int run = s_prefix + excl_in_block;
#pragma unroll
    for (int k = 0; k < kScanMaxChunk; ++k) {
        int idx = first + k;
        if (k < chunk && idx < T) {
            offsets[idx] = run;
            counts[(size_t)idx * kCounterStride] = run;   // the counter becomes the tile's emit cursor
            run += v[k];
        }
    }
    // statistics
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        lmax = max(lmax, __shfl_xor_sync(0xffffffffu, lmax, d));
        lbig += __shfl_xor_sync(0xffffffffu, lbig, d);
    }
    if (lane == 0) {
        if (lmax > 0) atomicMax(stats + 1, lmax);
        if (lbig > 0) atomicAdd(stats + 2, lbig);
    }
    if (blockIdx.x == gridDim.x - 1 && tid == 0) {
        int total = s_prefix + s_total;
        offsets[T] = total;
        stats[0] = total;
    }
}

// Bitonic sort of P (power of two) 64-bit keys held in `s` by the whole CTA.
template <typename Ptr>
__device__ __forceinline__ void bitonic_sort(Ptr s, int P, int nthreads) {
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < (P >> 1); i += nthreads) {
                int lo = ((i & ~(j - 1)) << 1) | (i & (j - 1));
                int hi = lo | j;
                bool asc = (lo & k) == 0;
                uint64_t a = s[lo], b = s[hi];
                if ((a > b) == asc) { s[lo] = b; s[hi] = a; }
            }
            __syncthreads();
        }
    }
}

// ---- register sort network of one warp -----------------------------------------------------------
// Each lane holds E consecutive keys (blocked layout, index = lane*E + slot) of a 32*E-key chunk.
// Bitonic stages with stride < E are compare-exchanges between a lane's own registers; strides >= E
// exchange whole registers with lane ^ (stride/E) through shuffles.  No barriers, no shared memory.
constexpr int kWarpSortMax = 512;
constexpr int kWarpSortSlice = kWarpSortMax + 32;   // u64 per warp incl. skew

__device__ __forceinline__ void cex(uint64_t& a, uint64_t& b, bool asc) {
    bool sw = (a > b) == asc;
    uint64_t lo = sw ? b : a, hi = sw ? a : b;
    a = lo; b = hi;
}

template <int E>
__device__ __forceinline__ void warp_sort_regs(uint64_t (&v)[E]) {
    constexpr int P = 32 * E;
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    // phase 1 (k <= E): every lane sorts its own E keys in registers, fully unrolled
#pragma unroll
    for (int k = 2; k <= E; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
            for (int s = 0; s < E; ++s) {
                if ((s & j) == 0) {
                    // direction = bit k of the global index lane*E + s
                    bool asc = (k < E) ? ((s & k) == 0) : ((lane & 1) == 0);
                    cex(v[s], v[s | j], asc);
                }
            }
        }
    }
    // phase 2 (k = 2E .. P): runtime loops keep the code small (the fully unrolled network
    // stalled on instruction fetch); each k = cross-lane stages by shuffle + an in-lane merge
#pragma unroll 1
    for (int k = 2 * E; k <= P; k <<= 1) {
        const bool asc = ((lane * E) & k) == 0;
#pragma unroll 1
        for (int lj = k / (2 * E); lj > 0; lj >>= 1) {
            const bool keep_min = ((lane & lj) == 0) == asc;
#pragma unroll
            for (int s = 0; s < E; ++s) {
                uint64_t o = __shfl_xor_sync(full, v[s], lj);
                uint64_t mn = v[s] < o ? v[s] : o, mx = v[s] < o ? o : v[s];
                v[s] = keep_min ? mn : mx;
            }
        }
#pragma unroll
        for (int j = E >> 1; j > 0; j >>= 1) {
#pragma unroll
            for (int s = 0; s < E; ++s)
                if ((s & j) == 0) cex(v[s], v[s | j], asc);
        }
    }
}

// ---- warp-per-tile sort (lists of up to 512 entries) --------------------------------------------
// Shared memory is used once to turn the coalesced load into the blocked layout (skewed by one slot
// per E to stay bank-conflict free) and once for the coalesced store.
template <int E>
__device__ __forceinline__ void warp_sort_tile(uint64_t* sw, const uint64_t* __restrict__ keys,
                                               int32_t* __restrict__ out, int start, int n) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    auto phys = [](int idx) { return idx + idx / E; };
#pragma unroll
    for (int s = 0; s < E; ++s) {
        int idx = lane + 32 * s;
        sw[phys(idx)] = (idx < n) ? keys[start + idx] : ~0ull;
    }
    __syncwarp(full);
    uint64_t v[E];
#pragma unroll
    for (int s = 0; s < E; ++s) v[s] = sw[phys(lane * E + s)];
    warp_sort_regs<E>(v);
    __syncwarp(full);
    uint32_t* s32 = reinterpret_cast<uint32_t*>(sw);
#pragma unroll
    for (int s = 0; s < E; ++s) s32[lane * E + s + lane] = (uint32_t)v[s];   // skew 1 word per lane
    __syncwarp(full);
#pragma unroll
    for (int s = 0; s < E; ++s) {
        int idx = lane + 32 * s;
        if (idx < n) out[start + idx] = (int32_t)s32[idx + idx / E];
    }
    __syncwarp(full);
}

__global__ void __launch_bounds__(256)
bin_sort_warp_kernel(int T, const int32_t* __restrict__ offsets, const uint64_t* __restrict__ keys,
                     int32_t* __restrict__ ids_sorted, int cap, int max_sorted) {
    __shared__ __align__(16) uint64_t s_keys[8 * kWarpSortSlice];
    const int tile = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (tile >= T) return;
    const int start = __ldg(offsets + tile);
    const int n = __ldg(offsets + tile + 1) - start;
    if (n <= 0 || start + n > cap) return;                         // cap: see ts_bin_emit
    if (n > max_sorted) {
        // longer than any size class this call launches (the host's bound was a guess): the list stays
        // unsorted, but it must hold VALID ids — the blend kernel queued behind us gathers through it
        // before the host has noticed and repeated the pass
        for (int i = threadIdx.x & 31; i < n; i += 32) ids_sorted[start + i] = (int32_t)(uint32_t)keys[start + i];
        return;
    }
    if (n > kWarpSortMax) return;
    uint64_t* sw = s_keys + (threadIdx.x >> 5) * kWarpSortSlice;
    if (n == 1) { if ((threadIdx.x & 31) == 0) ids_sorted[start] = (int32_t)(uint32_t)keys[start]; }
    else if (n <= 64) warp_sort_tile<2>(sw, keys, ids_sorted, start, n);
    else if (n <= 128) warp_sort_tile<4>(sw, keys, ids_sorted, start, n);
    else if (n <= 256) warp_sort_tile<8>(sw, keys, ids_sorted, start, n);
    else warp_sort_tile<16>(sw, keys, ids_sorted, start, n);
}

// ---- CTA-per-tile merge sort (lists of 513 .. 16384 entries) -------------------------------------
// NT threads sort P = NT*E keys.  Thread t owns positions [t*E, (t+1)*E) throughout: (1) each warp
// sorts its 32*E-key chunk with the register network above; (2) log2(NT/32) merge levels: the keys go
// to shared memory, every thread finds where ITS E outputs start in the two runs being merged (merge
// path: a binary search along the cross diagonal), merges E keys serially into registers, and the
// registers go back to shared memory for the next level.  O(n log n) compares and 2 barriers per
// level; the shared-memory bitonic network this replaces did log^2 n / 2 barrier-separated stages
// (0.056 -> 0.234 ms for 2x the pairs on the 2M scene, 0.39 ms on the dense 1M scene).
// Index skew (one slot per E) keeps the stride-E accesses of the owners bank-conflict free.
template <int E, int NT>
__device__ __forceinline__ void cta_merge_sort_tile(uint64_t* s, const uint64_t* __restrict__ keys,
                                                    int32_t* __restrict__ out, int start, int n) {
    constexpr int P = NT * E;
    const int tid = threadIdx.x;
    auto phys = [](int idx) { return idx + idx / E; };
    for (int i = tid; i < P; i += NT) s[phys(i)] = (i < n) ? keys[start + i] : ~0ull;
    __syncthreads();
    uint64_t v[E];
#pragma unroll
    for (int k = 0; k < E; ++k) v[k] = s[phys(tid * E + k)];
    warp_sort_regs<E>(v);
#pragma unroll 1
    for (int L = 32 * E; L < P; L <<= 1) {
        __syncthreads();                                   // the previous level's reads are done
#pragma unroll
        for (int k = 0; k < E; ++k) s[phys(tid * E + k)] = v[k];
        __syncthreads();
        const int o0 = tid * E;
        const int base = o0 & ~(2 * L - 1);                // the pair of runs [base, base+L) and [base+L, base+2L)
        const int d = o0 - base;                           // this thread's outputs start at rank d of the merged pair
        auto A = [&](int i) { return s[phys(base + i)]; };
        auto B = [&](int i) { return s[phys(base + L + i)]; };
        int lo = max(0, d - L), hi = min(d, L);
        while (lo < hi) {                                  // ties take A first (as the serial merge below)
            const int mid = (lo + hi) >> 1;
            if (A(mid) <= B(d - 1 - mid)) lo = mid + 1; else hi = mid;
        }
        int a = lo, b = d - lo;
        uint64_t ka = (a < L) ? A(a) : ~0ull, kb = (b < L) ? B(b) : ~0ull;
#pragma unroll
        for (int k = 0; k < E; ++k) {
            const bool take_a = (b >= L) || (a < L && ka <= kb);
            v[k] = take_a ? ka : kb;
            if (take_a) { ++a; ka = (a < L) ? A(a) : ~0ull; }
            else        { ++b; kb = (b < L) ? B(b) : ~0ull; }
        }
    }
    __syncthreads();
    uint32_t* s32 = reinterpret_cast<uint32_t*>(s);
#pragma unroll
    for (int k = 0; k < E; ++k) s32[tid * E + k + tid] = (uint32_t)v[k];     // skew 1 word per thread
    __syncthreads();
    for (int i = tid; i < n; i += NT) out[start + i] = (int32_t)s32[i + i / E];
}

// One CTA per tile; a launch handles the tiles with lo_count < n <= NT*E_HI (size classes keep the
// threads, shared memory and occupancy of a CTA matched to the list length).
template <int NT, int E_LO, int E_HI>
__global__ void __launch_bounds__(NT)
bin_sort_cta_kernel(int T, const int32_t* __restrict__ offsets, const uint64_t* __restrict__ keys,
                    int32_t* __restrict__ ids_sorted, int lo_count, int cap) {
    TS_DYN_SMEM(uint64_t, s_keys, 16);
    const int tile = blockIdx.x;
    const int start = __ldg(offsets + tile);
    const int n = __ldg(offsets + tile + 1) - start;
    if (n <= lo_count || n > NT * E_HI || start + n > cap) return;
    if (E_LO != E_HI && n <= NT * E_LO) cta_merge_sort_tile<E_LO, NT>(s_keys, keys, ids_sorted, start, n);
    else cta_merge_sort_tile<E_HI, NT>(s_keys, keys, ids_sorted, start, n);
}

constexpr size_t cta_sort_smem(int NT, int E) { return sizeof(uint64_t) * ((size_t)NT * E + NT); }

// Fallback for tiles whose list exceeds the shared-memory cap: same network, in global scratch.
__global__ void __launch_bounds__(1024)
bin_sort_big_kernel(int T, const int32_t* __restrict__ offsets, const uint64_t* __restrict__ keys,
                    int32_t* __restrict__ ids_sorted, int cap, int P, uint64_t* __restrict__ scratch,
                    int32_t* __restrict__ counter, int key_cap) {
    __shared__ int s_slot;
    const int tile = blockIdx.x;
    const int start = __ldg(offsets + tile);
    const int n = __ldg(offsets + tile + 1) - start;
    if (n <= cap || n > P || start + n > key_cap) return;
    if (threadIdx.x == 0) s_slot = atomicAdd(counter, 1);
    __syncthreads();
    uint64_t* buf = scratch + (size_t)s_slot * P;
    for (int i = threadIdx.x; i < P; i += 1024) buf[i] = (i < n) ? keys[start + i] : ~0ull;
    __syncthreads();
    bitonic_sort(buf, P, 1024);
    for (int i = threadIdx.x; i < n; i += 1024) ids_sorted[start + i] = (int32_t)(uint32_t)buf[i];
}

// emit advanced the cursors; a host that has to repeat emit (its key buffer was too small) restores them
__global__ void __launch_bounds__(kBinThreads)
bin_reset_cursors_kernel(int T, const int32_t* __restrict__ offsets, int32_t* __restrict__ cursors) {
    const int t = blockIdx.x * kBinThreads + threadIdx.x;
    if (t < T) cursors[(size_t)t * kCounterStride] = offsets[t];
}

// ---- launch order of the blend kernels: tiles by descending list length ---------------------------
// One CTA, counting sort on 256 length classes (class = min(n, 2047) / 8, longest first).  The blend
// grids are a few waves of CTAs whose cost is proportional to the list length; in raster order the
// last wave can hold long lists and the SMs that finish early idle (10 % of blend-backward's time at
// 1M Gaussians / 1080p).  Longest first, the tail is made of the cheapest tiles.
constexpr int kOrderThreads = 1024;
constexpr int kOrderClasses = 256;
__device__ __forceinline__ int order_class(int n) { return kOrderClasses - 1 - (min(n, 2047) >> 3); }

__global__ void __launch_bounds__(kOrderThreads)
bin_tile_order_kernel(int T, const int32_t* __restrict__ offsets, int32_t* __restrict__ order) {
    __shared__ int s_hist[kOrderClasses];
    __shared__ int s_base[kOrderClasses];
    const int tid = threadIdx.x;
    if (tid < kOrderClasses) s_hist[tid] = 0;
    __syncthreads();
    for (int t = tid; t < T; t += kOrderThreads)
        atomicAdd(&s_hist[order_class(__ldg(offsets + t + 1) - __ldg(offsets + t))], 1);
    __syncthreads();
    if (tid == 0) {
        int run = 0;
        for (int c = 0; c < kOrderClasses; ++c) { s_base[c] = run; run += s_hist[c]; }
    }
    __syncthreads();
    for (int t = tid; t < T; t += kOrderThreads) {
        const int c = order_class(__ldg(offsets + t + 1) - __ldg(offsets + t));
        order[atomicAdd(&s_base[c], 1)] = t;
    }
}

}  // namespace ts

#ifndef TS_HOST_EMU
extern "C" {

int ts_bin_smem_sort_cap(void) { return ts::kSmemSortCap; }

int ts_bin_count(int N, int CH, const float* xys, const float* depths, const int32_t* radii,
                 const float* conics, const float* opacity, const float* colors, int img_height,
                 int img_width, int tiles_x, int tiles_y, int cull_mode, int flags, float* recs,
                 int32_t* tile_counts, ts_stream_t stream) {
    (void)depths; (void)img_height; (void)img_width;
    if (N < 0 || CH < 1 || CH > 4 || tiles_x <= 0 || tiles_y <= 0) return TS_ERR_INVALID;
    const bool pack_only = (flags & TS_BIN_PACK_ONLY) != 0;
    if (!tile_counts && !pack_only) return TS_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    if (!pack_only)
        TS_CHECK_CUDA(cudaMemsetAsync(tile_counts, 0, sizeof(int32_t) * ts::kCounterStride * (size_t)tiles_x * tiles_y, st),
                      "ts_bin_count/memset");
    if (N == 0) return TS_OK;
    if (!xys || !radii || !conics || !opacity || !recs) return TS_ERR_INVALID;
    if (!ts::aligned16(recs) || (reinterpret_cast<uintptr_t>(xys) & 7u)) return TS_ERR_ALIGN;
    int grid = (N + ts::kBinThreads - 1) / ts::kBinThreads;
#define TS_LAUNCH_COUNT(C) \
    ts::bin_count_kernel<C><<<grid, ts::kBinThreads, 0, st>>>(N, (const float2*)xys, radii, conics, opacity, colors, tiles_x, tiles_y, cull_mode, flags, (float4*)recs, tile_counts)
    switch (CH) {
        case 1: TS_LAUNCH_COUNT(1); break;
        case 2: TS_LAUNCH_COUNT(2); break;
        case 3: TS_LAUNCH_COUNT(3); break;
        default: TS_LAUNCH_COUNT(4); break;
    }
#undef TS_LAUNCH_COUNT
    TS_CHECK_LAUNCH("ts_bin_count");
    return TS_OK;
}

int ts_bin_counter_stride(void) { return ts::kCounterStride; }

int ts_bin_scan_work_ints(void) { return ts::kScanWorkInts; }

int ts_bin_scan(int num_tiles, int32_t* tile_counts, int32_t* tile_offsets, int32_t* stats,
                int smem_sort_cap, ts_stream_t stream) {
    if (num_tiles <= 0 || !tile_counts || !tile_offsets || !stats) return TS_ERR_INVALID;
    // chunk tiles per thread so that the grid stays within kScanMaxBlocks co-resident CTAs
    int chunk = (num_tiles + 1024 * ts::kScanMaxBlocks - 1) / (1024 * ts::kScanMaxBlocks);
    if (chunk < 1) chunk = 1;
    if (chunk > ts::kScanMaxChunk) return TS_ERR_CAPACITY;    // > 2M tiles (a > 500 Mpixel image)
    int grid = (num_tiles + 1024 * chunk - 1) / (1024 * chunk);
    cudaStream_t st = (cudaStream_t)stream;
    TS_CHECK_CUDA(cudaMemsetAsync(stats, 0, sizeof(int32_t) * ts::kScanWorkInts, st), "ts_bin_scan/memset");
    ts::bin_scan_kernel<<<grid, 1024, 0, st>>>(num_tiles, chunk, tile_counts, tile_offsets, stats, smem_sort_cap);
    TS_CHECK_LAUNCH("ts_bin_scan");
    return TS_OK;
}

int ts_bin_emit(int N, const float* depths, const int32_t* radii, const float* recs, int tiles_x,
                int tiles_y, int cull_mode, const int32_t* tile_offsets, int32_t* cursors,
                uint64_t* keys, int capacity, ts_stream_t stream) {
    if (N < 0 || tiles_x <= 0 || tiles_y <= 0) return TS_ERR_INVALID;
    if (N == 0) return TS_OK;
    if (!depths || !radii || !recs || !tile_offsets || !cursors || !keys) return TS_ERR_INVALID;
    if (!ts::aligned16(recs)) return TS_ERR_ALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    int grid = (N + ts::kBinThreads - 1) / ts::kBinThreads;
    ts::bin_emit_kernel<<<grid, ts::kBinThreads, 0, st>>>(N, depths, radii, (const float4*)recs, tiles_x,
                                                          tiles_y, cull_mode, cursors, keys,
                                                          capacity > 0 ? capacity : INT32_MAX);
    TS_CHECK_LAUNCH("ts_bin_emit");
    return TS_OK;
}

int ts_bin_reset_cursors(int num_tiles, const int32_t* tile_offsets, int32_t* cursors, ts_stream_t stream) {
    if (num_tiles <= 0 || !tile_offsets || !cursors) return TS_ERR_INVALID;
    ts::bin_reset_cursors_kernel<<<(num_tiles + ts::kBinThreads - 1) / ts::kBinThreads, ts::kBinThreads, 0, (cudaStream_t)stream>>>(num_tiles, tile_offsets, cursors);
    TS_CHECK_LAUNCH("ts_bin_reset_cursors");
    return TS_OK;
}

int ts_bin_tile_order(int num_tiles, const int32_t* tile_offsets, int32_t* tile_order, ts_stream_t stream) {
    if (num_tiles <= 0 || !tile_offsets || !tile_order) return TS_ERR_INVALID;
    ts::bin_tile_order_kernel<<<1, ts::kOrderThreads, 0, (cudaStream_t)stream>>>(num_tiles, tile_offsets, tile_order);
    TS_CHECK_LAUNCH("ts_bin_tile_order");
    return TS_OK;
}

int ts_bin_sort(int num_tiles, const int32_t* tile_offsets, uint64_t* keys, int32_t* ids_sorted,
                int max_count, int n_big_tiles, uint64_t* big_scratch, int32_t* big_counter,
                int capacity, ts_stream_t stream) {
    if (num_tiles <= 0 || !tile_offsets) return TS_ERR_INVALID;
    if (max_count <= 0) return TS_OK;
    if (!keys || !ids_sorted) return TS_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    const int cap = capacity > 0 ? capacity : INT32_MAX;
    // lists of <= 512 entries: one warp per tile, keys in registers
    ts::bin_sort_warp_kernel<<<(num_tiles + 7) / 8, 256, 0, st>>>(num_tiles, tile_offsets, keys, ids_sorted, cap,
                                                                  n_big_tiles > 0 ? INT32_MAX : min(max_count, ts::kSmemSortCap));
    TS_CHECK_LAUNCH("ts_bin_sort/warp");
    // longer lists: one CTA per tile (register chunk sort + merge-path levels); a class is launched
    // only when max_count says some tile can need it
    if (max_count > ts::kWarpSortMax) {
        constexpr size_t smem = ts::cta_sort_smem(256, 8);
        ts::bin_sort_cta_kernel<256, 4, 8><<<num_tiles, 256, smem, st>>>(num_tiles, tile_offsets, keys, ids_sorted,
                                                                       ts::kWarpSortMax, cap);
        TS_CHECK_LAUNCH("ts_bin_sort/cta256");
    }
    if (max_count > 2048) {
        constexpr size_t smem = ts::cta_sort_smem(512, 16);
        // per device and context, cheap: set on every call (a process may drive several GPUs)
        TS_CHECK_CUDA(cudaFuncSetAttribute(ts::bin_sort_cta_kernel<512, 8, 16>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "ts_bin_sort/attr");
        ts::bin_sort_cta_kernel<512, 8, 16><<<num_tiles, 512, smem, st>>>(num_tiles, tile_offsets, keys, ids_sorted,
                                                                        2048, cap);
        TS_CHECK_LAUNCH("ts_bin_sort/cta512");
    }
    if (max_count > 8192) {
        constexpr size_t smem = ts::cta_sort_smem(1024, 16);
        TS_CHECK_CUDA(cudaFuncSetAttribute(ts::bin_sort_cta_kernel<1024, 16, 16>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "ts_bin_sort/attr");
        ts::bin_sort_cta_kernel<1024, 16, 16><<<num_tiles, 1024, smem, st>>>(num_tiles, tile_offsets, keys, ids_sorted,
                                                                           8192, cap);
        TS_CHECK_LAUNCH("ts_bin_sort/cta1024");
    }
    if (n_big_tiles > 0) {
        if (!big_scratch || !big_counter) return TS_ERR_CAPACITY;
        int P = 2;
        while (P < max_count) P <<= 1;
        TS_CHECK_CUDA(cudaMemsetAsync(big_counter, 0, sizeof(int32_t), st), "ts_bin_sort/memset");
        ts::bin_sort_big_kernel<<<num_tiles, 1024, 0, st>>>(num_tiles, tile_offsets, keys, ids_sorted,
                                                            ts::kSmemSortCap, P, big_scratch, big_counter, cap);
        TS_CHECK_LAUNCH("ts_bin_sort/big");
    }
    return TS_OK;
}

}  // extern "C"
#endif  // !TS_HOST_EMU
