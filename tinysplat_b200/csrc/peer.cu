// SURVEY 8(e) — data-parallel gradient exchange done by the kernels themselves over NVLink peer
// memory (no NCCL call inside the step).  [The reference renders one camera per step on one device,
// REF scripts/train.py:54-55; the camera-sharded batch is this repo's addition.]
//
// Rank-structured exchange.  After blend-backward every rank holds, for ITS view, one packed
// 48-byte gradient row per Gaussian.  Everything downstream is a per-Gaussian function of those rows
// and the views' cameras, so instead of all-reducing the finished 236-byte gradients:
//   * geometry (S-sums, v_opacity, depth cotangent: 32 B) goes only to the rank that OWNS the
//     Gaussian (a shard of N/world rows): projection-backward is linear in nothing, it has to be
//     evaluated per view, and the owner does that for all views (ts_project_bwd_views) and stores
//     the finished 44 B of mean/scale/quat/opacity gradient straight into EVERY rank's buffer;
//   * colour cotangents (12 B) go to EVERY rank: the SH gradient is a sum over views of
//     basis(dir_view) x v_rgb_view, rank <= world, so shipping the factors (12 B per view) and
//     rebuilding the 192-byte SH gradient locally moves half of what shipping the product would.
// Per rank and Gaussian at 8 ranks: 28 + 84 + 38.5 = 150 B over NVLink instead of the 413 B a ring
// all-reduce of 236 B moves; at 2 ranks 50 B instead of 236 B.
//
// Kernels here: ts_dp_push (cleans the rows like ts_dp_prepare and pushes them with coalesced
// 128-bit stores to peer-mapped addresses), ts_peer_barrier (system-scope release/acquire flags in
// peer memory).  Peer buffers are plain cudaMalloc allocations shared through CUDA IPC handles.
#include <cstdio>
#include <cstdlib>
#include "ts_common.cuh"
#include "ts_peer.cuh"

namespace ts {

constexpr int kPushThreads = 128;      // threads per CTA
constexpr int kPushRows = 256;         // rows per block of work (two per thread): one owner, one set of bulk stores
constexpr int kPushCtasPerSm = 4;      // resident footprint of the push grid: 512 threads, 44 KB of shared memory per SM

// Persistent: a grid of at most (#SMs x kPushCtasPerSm) CTAs walks the 256-row blocks.  The kernel is
// NVLink-bound and its CTAs spend their life waiting for remote stores; a grid that fills the SMs with
// such waiters (the first version: 8 x 256 threads per SM) keeps the kernels of the other streams —
// the shard backward of the PREVIOUS piece — off the machine, and the pipeline of ts_dp_exchange_peer
// degenerates into a sequence.  Here a CTA keeps its stores in flight across iterations (it waits only
// until the copy engine has READ the shared buffer before refilling it, and for completion once, at the end).
__global__ void __launch_bounds__(kPushThreads)
dp_push_kernel(int N, int Ns, int Npad, int world, int rank, const int32_t* __restrict__ radii,
               const uint8_t* __restrict__ clamp_mask, const float4* __restrict__ recs,
               const float4* __restrict__ grads, const float* __restrict__ cam_row, PeerPtrs geo,
               PeerPtrs rgb, PeerPtrs cams, float2* __restrict__ v_xys, int what) {
    // what: bit 0 = geometry rows (+ this view's d loss / d xy and its camera), bit 1 = colour rows.  The
    // exchange pushes the geometry first and signals it separately: the owners' projection-backward then
    // runs under the (three times larger) colour transfer (ts_dp_exchange_peer).
    const bool do_geo = (what & 1) != 0, do_rgb = (what & 2) != 0;
    __shared__ __align__(128) float4 s_geo[2][kPushRows * 2];     // the block's geometry rows, 2 x 8 KB
    __shared__ __align__(128) float s_rgb[2][kPushRows * 3];      // the block's colour rows, 2 x 3 KB
    const int tid = threadIdx.x;
    if (blockIdx.x == 0 && do_geo) {                // this view's camera -> every rank's cams[rank]
        for (int idx = tid; idx < kCamRowFloats * world; idx += kPushThreads) {
            const int r = idx / kCamRowFloats, k = idx - r * kCamRowFloats;
            (reinterpret_cast<float*>(cams.p[r]) + (size_t)rank * kCamRowFloats)[k] = __ldg(cam_row + k);
        }
    }
    const int nblocks = (N + kPushRows - 1) / kPushRows;
    // every rank starts with the blocks its NEXT neighbour owns: at any moment the ranks' geometry rows
    // go to different owners (in plain block order all of them would hit owner 0, then owner 1, ...)
    const int rot = (int)(((long long)((rank + 1) % world) * Ns) / kPushRows) % max(nblocks, 1);
    int it = 0;
    for (int b0 = blockIdx.x; b0 < nblocks; b0 += gridDim.x, ++it) {
        const int buf = it & 1;
        const int blk = b0 + rot < nblocks ? b0 + rot : b0 + rot - nblocks;
        const int item0 = blk * kPushRows;
        // the bulk stores issued two iterations ago read this buffer: wait until at most ONE group
        // (the previous iteration's) is still reading
        if (it >= 2) {
            if (tid == 0) bulk_wait_read1();
            __syncthreads();
        }
#pragma unroll
        for (int h = 0; h < kPushRows / kPushThreads; ++h) {
            const int t = h * kPushThreads + tid;
            const int i = item0 + t;
            float4 g0 = make_float4(0.f, 0.f, 0.f, 0.f), g1 = g0, g2 = g0;
            if (i < N) {
                const bool live = __ldg(radii + i) > 0;     // culled in this view: an exact zero row
                if (live) {
                    g0 = __ldg(grads + 3 * (size_t)i);
                    g1 = __ldg(grads + 3 * (size_t)i + 1);
                    g2 = __ldg(grads + 3 * (size_t)i + 2);
                    if (clamp_mask) {                       // SH clamp(rgb + 0.5, min=0) [REF rasterize.py:39]
                        const unsigned m = clamp_mask[i];
                        if (!(m & 1u)) g2.x = 0.f;
                        if (!(m & 2u)) g2.y = 0.f;
                        if (!(m & 4u)) g2.z = 0.f;
                    }
                }
                if (v_xys && do_geo) {                      // this view's d loss / d xy (densification statistic)
                    float2 v = make_float2(0.f, 0.f);
                    if (live) {
                        const float4 q1 = __ldg(recs + 3 * (size_t)i + 1);   // {.5 log2e a, log2e b, .5 log2e c, opacity}
                        const float a = q1.x * (2.f / kLog2e), b = q1.y * (1.f / kLog2e), c = q1.z * (2.f / kLog2e);
                        v = make_float2(a * g0.x + b * g0.y, b * g0.x + c * g0.y);
                    }
                    v_xys[i] = v;
                }
            }
            s_geo[buf][2 * t] = g0;
            s_geo[buf][2 * t + 1] = make_float4(g1.x, g1.y, g2.w, 0.f);
            s_rgb[buf][3 * t] = g2.x; s_rgb[buf][3 * t + 1] = g2.y; s_rgb[buf][3 * t + 2] = g2.z;
        }
        const int nvalid = min(kPushRows, N - item0);
        const size_t rgb_off = (size_t)rank * Npad * 3 + (size_t)item0 * 3;      // floats, in every rank's rgb buffer
        const int owner0 = item0 / Ns;
        // A full block whose rows have ONE owner (always, when shard_rows is a multiple of 256) ships as
        // TMA bulk stores: one 8 KB copy of geometry rows to the owner, one 3 KB copy of colour rows per
        // rank, issued by one thread — the copy engine generates the NVLink traffic, not 10 stores per lane.
        const bool bulk = nvalid == kPushRows && owner0 == (item0 + kPushRows - 1) / Ns;
        if (bulk) {
            fence_proxy_async();        // generic-proxy smem writes -> visible to the copy engine
            __syncthreads();
            if (tid == 0) {
                if (do_geo) {
                    float4* gdst = reinterpret_cast<float4*>(geo.p[owner0]) + ((size_t)rank * Ns + (item0 - owner0 * Ns)) * 2;
                    bulk_s2g(gdst, s_geo[buf], kPushRows * 32);
                }
                for (int d = 0; d < world && do_rgb; ++d) {     // start at the next rank: spreads the links
                    const int r = (rank + 1 + d) % world;
                    bulk_s2g(reinterpret_cast<float*>(rgb.p[r]) + rgb_off, s_rgb[buf], kPushRows * 12);
                }
            }
            if (tid == 0) bulk_commit();                // one group per iteration (empty groups complete at once)
        } else {
            __syncthreads();
            for (int t = tid; t < nvalid && do_geo; t += kPushThreads) {
                const int i = item0 + t;
                const int owner = i / Ns, il = i - owner * Ns;
                float4* dst = reinterpret_cast<float4*>(geo.p[owner]) + ((size_t)rank * Ns + il) * 2;
                dst[0] = s_geo[buf][2 * t];
                dst[1] = s_geo[buf][2 * t + 1];
            }
            const int nfl = do_rgb ? nvalid * 3 : 0;
            const int nv4 = nfl >> 2;
            for (int idx = tid; idx < nv4 * world; idx += kPushThreads) {
                const int d = idx / nv4, k = idx - d * nv4;
                const int r = (rank + 1 + d) % world;
                reinterpret_cast<float4*>(reinterpret_cast<float*>(rgb.p[r]) + rgb_off)[k] = reinterpret_cast<const float4*>(s_rgb[buf])[k];
            }
            const int ntail = nfl - (nv4 << 2);
            for (int idx = tid; idx < ntail * world; idx += kPushThreads) {
                const int d = idx / ntail, k = (nv4 << 2) + idx - d * ntail;
                const int r = (rank + 1 + d) % world;
                (reinterpret_cast<float*>(rgb.p[r]) + rgb_off)[k] = s_rgb[buf][k];
            }
            if (tid == 0) bulk_commit();
            __syncthreads();            // the generic stores above read the buffer: done before it is refilled
        }
    }
    // performed, not just read: the flag signal that follows the kernel publishes them
    if (tid == 0) bulk_wait0();
}

static int g_exchange_split = -1;
static bool exchange_split() {
    if (g_exchange_split < 0) {
#ifndef TS_HOST_EMU
        const char* e = getenv("TINYSPLAT_B200_PEER_SPLIT");
        g_exchange_split = (e && e[0] == '0') ? 0 : 1;
#else
        g_exchange_split = 1;
#endif
    }
    return g_exchange_split == 1;
}

// Flags live one per 128-byte line: flags[slot][source rank][32 words].
// mode bit 0: SIGNAL (tell every rank that everything this stream did before has happened);
// mode bit 1: WAIT (until every rank has signalled `epoch` in this slot).  Signal and wait may sit on
// different streams: the pushing stream only signals and never blocks on a peer, the consuming stream waits.
__global__ void peer_barrier_kernel(int world, int rank, PeerPtrs flags, int slot, uint32_t epoch,
                                    uint32_t* __restrict__ err, unsigned long long timeout_ns, int mode) {
    const int t = threadIdx.x;
    if (t >= world) return;
#ifndef TS_HOST_EMU
    if (mode & 1) {
        __threadfence_system();     // everything this stream wrote before (also to peers) is ordered before the signal
        uint32_t* remote = reinterpret_cast<uint32_t*>(flags.p[t]) + ((size_t)slot * kMaxPeers + rank) * 32;
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(epoch) : "memory");
    }
    if (!(mode & 2)) return;
    const uint32_t* local = reinterpret_cast<const uint32_t*>(flags.p[rank]) + ((size_t)slot * kMaxPeers + t) * 32;
    unsigned long long t0, now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        uint32_t v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(local) : "memory");
        if ((int32_t)(v - epoch) >= 0) break;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (now - t0 > timeout_ns) {    // a peer never arrived: report instead of hanging the GPU
            atomicExch(err, 1u + (uint32_t)t);
            break;
        }
        __nanosleep(200);
    }
    __threadfence_system();
#else
    (void)flags; (void)slot; (void)epoch; (void)err; (void)timeout_ns; (void)rank; (void)mode;
#endif
}

}  // namespace ts

#ifndef TS_HOST_EMU
extern "C" {

int ts_peer_max_ranks(void) { return ts::kMaxPeers; }
int ts_peer_ipc_handle_bytes(void) { return (int)sizeof(cudaIpcMemHandle_t); }

int ts_peer_alloc(int64_t bytes, void** dev_ptr_host) {
    if (bytes <= 0 || !dev_ptr_host) return TS_ERR_INVALID;
    void* p = nullptr;
    TS_CHECK_CUDA(cudaMalloc(&p, (size_t)bytes), "ts_peer_alloc/cudaMalloc");
    cudaError_t e = cudaMemset(p, 0, (size_t)bytes);
    if (e != cudaSuccess) { cudaFree(p); ts::set_last_error("ts_peer_alloc/memset", e); return TS_ERR_CUDA; }
    *dev_ptr_host = p;
    return TS_OK;
}

int ts_peer_free(void* dev_ptr) {
    if (!dev_ptr) return TS_OK;
    TS_CHECK_CUDA(cudaFree(dev_ptr), "ts_peer_free");
    return TS_OK;
}

int ts_peer_ipc_get(void* dev_ptr, void* handle_host) {
    if (!dev_ptr || !handle_host) return TS_ERR_INVALID;
    cudaIpcMemHandle_t h;
    TS_CHECK_CUDA(cudaIpcGetMemHandle(&h, dev_ptr), "ts_peer_ipc_get");
    memcpy(handle_host, &h, sizeof(h));
    return TS_OK;
}

int ts_peer_ipc_open(const void* handle_host, void** dev_ptr_host) {
    if (!handle_host || !dev_ptr_host) return TS_ERR_INVALID;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle_host, sizeof(h));
    void* p = nullptr;
    TS_CHECK_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess), "ts_peer_ipc_open");
    *dev_ptr_host = p;
    return TS_OK;
}

int ts_peer_ipc_close(void* dev_ptr) {
    if (!dev_ptr) return TS_OK;
    TS_CHECK_CUDA(cudaIpcCloseMemHandle(dev_ptr), "ts_peer_ipc_close");
    return TS_OK;
}

int ts_dp_push(int N, int shard_rows, int padded_rows, int world, int rank, const int32_t* radii,
               const uint8_t* clamp_mask, const float* recs, const float* grads, const float* cam_row,
               void* const* geo_ptrs_host, void* const* rgb_ptrs_host, void* const* cam_ptrs_host,
               float* v_xys, int what, ts_stream_t stream) {
    if (what < 1 || what > 3) return TS_ERR_INVALID;
    if (N < 0 || world < 1 || world > ts::kMaxPeers || rank < 0 || rank >= world || shard_rows <= 0 ||
        (shard_rows % 4) != 0 || padded_rows < N || (padded_rows % 4) != 0 ||
        (int64_t)shard_rows * world < N)
        return TS_ERR_INVALID;
    if (!geo_ptrs_host || !rgb_ptrs_host || !cam_ptrs_host || !cam_row) return TS_ERR_INVALID;
    ts::PeerPtrs geo{}, rgb{}, cams{};
    for (int r = 0; r < world; ++r) {
        geo.p[r] = geo_ptrs_host[r]; rgb.p[r] = rgb_ptrs_host[r]; cams.p[r] = cam_ptrs_host[r];
        if (!geo.p[r] || !rgb.p[r] || !cams.p[r]) return TS_ERR_INVALID;
        if (!ts::aligned16(geo.p[r]) || !ts::aligned16(rgb.p[r])) return TS_ERR_ALIGN;
    }
    if (N > 0 && (!radii || !recs || !grads)) return TS_ERR_INVALID;
    if (N > 0 && (!ts::aligned16(recs) || !ts::aligned16(grads) || (v_xys && (reinterpret_cast<uintptr_t>(v_xys) & 7u))))
        return TS_ERR_ALIGN;
    int dev = 0, sms = 148;
    TS_CHECK_CUDA(cudaGetDevice(&dev), "ts_dp_push/device");
    TS_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev), "ts_dp_push/attr");
    const int nblocks = (N + ts::kPushRows - 1) / ts::kPushRows;
    const int grid = max(1, min(nblocks, sms * ts::kPushCtasPerSm));               // block 0 always ships the camera
    ts::dp_push_kernel<<<grid, ts::kPushThreads, 0, (cudaStream_t)stream>>>(
        N, shard_rows, padded_rows, world, rank, radii, clamp_mask, (const float4*)recs, (const float4*)grads,
        cam_row, geo, rgb, cams, (float2*)v_xys, what);
    TS_CHECK_LAUNCH("ts_dp_push");
    return TS_OK;
}

int ts_peer_barrier(int world, int rank, void* const* flag_ptrs_host, int slot, uint32_t epoch,
                    uint32_t* err_flag, double timeout_s, int mode, ts_stream_t stream) {
    if (world < 1 || world > ts::kMaxPeers || rank < 0 || rank >= world || slot < 0 || slot >= ts::kBarrierSlots ||
        !flag_ptrs_host || !err_flag || mode < 1 || mode > 3)
        return TS_ERR_INVALID;
    ts::PeerPtrs flags{};
    for (int r = 0; r < world; ++r) {
        flags.p[r] = flag_ptrs_host[r];
        if (!flags.p[r]) return TS_ERR_INVALID;
    }
    if (!(timeout_s > 0)) timeout_s = 10.0;
    ts::peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(world, rank, flags, slot, epoch, err_flag,
                                                                 (unsigned long long)(timeout_s * 1e9), mode);
    TS_CHECK_LAUNCH("ts_peer_barrier");
    return TS_OK;
}

int ts_peer_barrier_slots(void) { return ts::kBarrierSlots; }

// -1: TINYSPLAT_B200_PEER_SPLIT environment variable / default (1); 0: one push per piece; 1: geometry and
// colour rows pushed and signalled separately (see ts_dp_exchange_peer)
int ts_dp_exchange_split(int mode) {
    if (mode < -1 || mode > 1) return TS_ERR_INVALID;
    ts::g_exchange_split = mode;
    return TS_OK;
}

// ---- the whole data-parallel backward tail as ONE call --------------------------------------------
// A three-stream pipeline over row pieces (see the header).  Everything here is host-side launch
// logic: the ~30 launches, event records and stream waits of a 4-piece exchange cost ~0.1 ms issued
// from here and ~0.4 ms issued one by one from Python — more than blend-backward leaves the CPU.
namespace {
struct EventPool {          // per device; events are reused round-robin (disable-timing: cheap to record)
    static constexpr int kN = 64;
    cudaEvent_t ev[kN];
    int next = 0;
    bool ready = false;
};
EventPool g_pools[16];

// Optional timeline of one exchange (ts_dp_exchange_timeline): timing events recorded between the
// launches on each stream; read back by ts_dp_exchange_timeline_read after a device synchronize.
struct Timeline {
    static constexpr int kMax = 64;
    bool on = false, ready = false;
    int n = 0;
    cudaEvent_t ev[kMax];
    char label[kMax][24];
};
Timeline g_tl;

void tl_mark(cudaStream_t st, const char* what, int piece) {
    if (!g_tl.on || g_tl.n >= Timeline::kMax) return;
    if (!g_tl.ready) {
        for (int i = 0; i < Timeline::kMax; ++i) cudaEventCreate(&g_tl.ev[i]);
        g_tl.ready = true;
    }
    snprintf(g_tl.label[g_tl.n], sizeof(g_tl.label[0]), "%s%d", what, piece);
    cudaEventRecord(g_tl.ev[g_tl.n++], st);
}

cudaEvent_t pool_event(int dev) {
    EventPool& P = g_pools[dev & 15];
    if (!P.ready) {
        for (int i = 0; i < EventPool::kN; ++i) cudaEventCreateWithFlags(&P.ev[i], cudaEventDisableTiming);
        P.ready = true;
    }
    cudaEvent_t e = P.ev[P.next];
    P.next = (P.next + 1) % EventPool::kN;
    return e;
}
}  // namespace

// Debug: ts_dp_exchange_timeline(1) makes every following ts_dp_exchange_peer record timing events
// between its launches; after a device synchronize ts_dp_exchange_timeline_read returns the number of
// marks of the LAST exchange and writes "label ms-since-start" lines into buf.
int ts_dp_exchange_timeline(int enable) { g_tl.on = enable != 0; return TS_OK; }
int ts_dp_exchange_timeline_read(char* buf, int buf_bytes) {
    if (!buf || buf_bytes <= 0) return TS_ERR_INVALID;
    int used = 0;
    buf[0] = 0;
    for (int i = 1; i < g_tl.n; ++i) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, g_tl.ev[0], g_tl.ev[i]) != cudaSuccess) { cudaGetLastError(); continue; }
        int w = snprintf(buf + used, (size_t)(buf_bytes - used), "%s %.4f\n", g_tl.label[i], ms);
        if (w < 0 || w >= buf_bytes - used) break;
        used += w;
    }
    return g_tl.n;
}

int ts_dp_exchange_peer(int N, int K, int degree, int world, int rank, int n_pieces, const int32_t* piece_plan_host,
                        int padded_rows, const int32_t* radii, const uint8_t* clamp_mask, const float* recs,
                        const float* grads, const float* cam_row, const float* means3d, const float* scales,
                        const float* quats, const float* opacity_logits, void* const* peer_bases_host,
                        const int64_t* seg_offsets_host, int img_height, int img_width, int proj_flags,
                        float out_scale, uint32_t epoch, double timeout_s, float* v_xys, ts_stream_t main_stream,
                        ts_stream_t side_stream, ts_stream_t side2_stream) {
    const bool split = ts::exchange_split();
    if (N < 0 || world < 1 || world > ts::kMaxPeers || rank < 0 || rank >= world || n_pieces < 0 ||
        (split ? 2 : 1) * n_pieces >= ts::kBarrierSlots || !peer_bases_host || !seg_offsets_host || !cam_row || (n_pieces > 0 && !piece_plan_host))
        return TS_ERR_INVALID;
    enum { SEG_FLAGS, SEG_ERR, SEG_CAMS, SEG_GEO, SEG_RGB, SEG_REST, SEG_DC, SEG_MEANS, SEG_SCALES, SEG_QUATS, SEG_LOGIT };
    cudaStream_t sm = (cudaStream_t)main_stream, s1 = (cudaStream_t)side_stream, s2 = (cudaStream_t)side2_stream;
    int dev = 0;
    TS_CHECK_CUDA(cudaGetDevice(&dev), "ts_dp_exchange_peer/device");
    auto seg = [&](int r, int which, int64_t byte_off) -> void* {
        return (void*)((char*)peer_bases_host[r] + seg_offsets_host[which] + byte_off);
    };
    auto table = [&](void** t, int which, int64_t byte_off) { for (int r = 0; r < world; ++r) t[r] = seg(r, which, byte_off); };
    void *t_flags[ts::kMaxPeers], *t_geo[ts::kMaxPeers], *t_rgb[ts::kMaxPeers], *t_cams[ts::kMaxPeers];
    void *t_m[ts::kMaxPeers], *t_s[ts::kMaxPeers], *t_q[ts::kMaxPeers], *t_l[ts::kMaxPeers];
    table(t_flags, SEG_FLAGS, 0);
    table(t_cams, SEG_CAMS, 0);
    uint32_t* err = (uint32_t*)seg(rank, SEG_ERR, 0);
    const float* cams_local = (const float*)seg(rank, SEG_CAMS, 0);
    const int R = (K - 1) * 3;
    int rc;
    g_tl.n = 0;
    tl_mark(sm, "start", 0);
    // whoever read last step's gradients (views of my segment) on the main stream is done
    cudaEvent_t e0 = pool_event(dev);
    TS_CHECK_CUDA(cudaEventRecord(e0, sm), "ts_dp_exchange_peer/event");
    if (s1 != sm) TS_CHECK_CUDA(cudaStreamWaitEvent(s1, e0, 0), "ts_dp_exchange_peer/wait");
    if (s2 != sm && s2 != s1) TS_CHECK_CUDA(cudaStreamWaitEvent(s2, e0, 0), "ts_dp_exchange_peer/wait");
    if (n_pieces == 0) {    // nothing to push, but the camera and the barrier keep the ranks in step
        table(t_geo, SEG_GEO, 0);
        table(t_rgb, SEG_RGB, 0);
        rc = ts_dp_push(0, 4, padded_rows > 0 ? padded_rows : 4 * world, world, rank, nullptr, nullptr, nullptr, nullptr,
                        cam_row, t_geo, t_rgb, t_cams, nullptr, 3, main_stream);
        if (rc != TS_OK) return rc;
    }
    for (int c = 0; c < n_pieces; ++c) {
        const int64_t r0 = piece_plan_host[4 * c], n = piece_plan_host[4 * c + 1], ns_c = piece_plan_host[4 * c + 2],
                      g0 = piece_plan_host[4 * c + 3];
        if (r0 < 0 || n <= 0 || r0 + n > N || ns_c <= 0) return TS_ERR_INVALID;
        const int64_t geo_off = (int64_t)world * g0 * 32;
        table(t_geo, SEG_GEO, geo_off);
        table(t_rgb, SEG_RGB, 12 * r0);
        const int64_t s0 = r0 + (int64_t)rank * ns_c;      // my shard of the piece (global rows)
        const int64_t ns = (int64_t)(rank + 1) * ns_c < n ? ns_c : n - (int64_t)rank * ns_c;
        auto push = [&](int what) {
            return ts_dp_push((int)n, (int)ns_c, padded_rows, world, rank, radii + r0, clamp_mask ? clamp_mask + r0 : nullptr,
                              recs + 12 * r0, grads + 12 * r0, cam_row, t_geo, t_rgb, t_cams,
                              v_xys ? v_xys + 2 * r0 : nullptr, what, main_stream);
        };
        auto shard_projection = [&]() -> int {             // over all views for MY shard, stored into every rank
            if (ns <= 0) return TS_OK;
            table(t_m, SEG_MEANS, 12 * s0);
            table(t_s, SEG_SCALES, 12 * s0);
            table(t_q, SEG_QUATS, 16 * s0);
            table(t_l, SEG_LOGIT, 4 * s0);
            int r = ts_project_bwd_views_peer(world, (int)ns, means3d + 3 * s0, scales + 3 * s0, 1.0f, quats + 4 * s0,
                                              cams_local, img_height, img_width, proj_flags,
                                              (const float*)seg(rank, SEG_GEO, geo_off), ns_c * 8,
                                              opacity_logits ? opacity_logits + s0 : nullptr, out_scale, world,
                                              (rank + 1) % world, t_m, t_s, t_q, t_l, side2_stream);
            if (r == TS_OK) tl_mark(s2, "proj_done", c);
            return r;
        };
        auto sh_gradient = [&]() -> int {                   // of the piece's Gaussians, from all views' colours (local)
            int r = ts_sh_bwd_views_rgb(world, (int)n, degree, K, means3d + 3 * r0, cams_local,
                                        (const float*)seg(rank, SEG_RGB, 12 * r0), (int64_t)padded_rows * 3, out_scale,
                                        (float*)seg(rank, SEG_DC, 12 * r0), (float*)seg(rank, SEG_REST, 4 * (int64_t)R * r0),
                                        side_stream);
            if (r == TS_OK) tl_mark(s1, "sh_done", c);
            return r;
        };
        if (split) {
            // main: geometry rows (28 of the 112 remote MB per rank at 8 GPUs) first, signalled on their own:
            // the owners' projection-backward (issue-bound) then runs under the colour transfer
            const int slot_geo = 2 * c, slot_rgb = 2 * c + 1;
            if ((rc = push(1)) != TS_OK) return rc;
            tl_mark(sm, "geo_pushed", c);
            if ((rc = ts_peer_barrier(world, rank, t_flags, slot_geo, epoch, err, timeout_s, 1, main_stream)) != TS_OK) return rc;
            if ((rc = push(2)) != TS_OK) return rc;
            tl_mark(sm, "rgb_pushed", c);
            if ((rc = ts_peer_barrier(world, rank, t_flags, slot_rgb, epoch, err, timeout_s, 1, main_stream)) != TS_OK) return rc;
            if ((rc = ts_peer_barrier(world, rank, t_flags, slot_geo, epoch, err, timeout_s, 2, side2_stream)) != TS_OK) return rc;
            tl_mark(s2, "geo_landed", c);
            if ((rc = shard_projection()) != TS_OK) return rc;
            if ((rc = ts_peer_barrier(world, rank, t_flags, slot_rgb, epoch, err, timeout_s, 2, side_stream)) != TS_OK) return rc;
            tl_mark(s1, "rgb_landed", c);
            if ((rc = sh_gradient()) != TS_OK) return rc;
        } else {
            // main: push the piece, signal; never waits for a peer
            if ((rc = push(3)) != TS_OK) return rc;
            tl_mark(sm, "pushed", c);
            if ((rc = ts_peer_barrier(world, rank, t_flags, c, epoch, err, timeout_s, 1, main_stream)) != TS_OK) return rc;
            // side: every rank's rows of this piece have landed here -> SH gradient of the piece (local)
            if ((rc = ts_peer_barrier(world, rank, t_flags, c, epoch, err, timeout_s, 2, side_stream)) != TS_OK) return rc;
            cudaEvent_t landed = pool_event(dev);
            TS_CHECK_CUDA(cudaEventRecord(landed, s1), "ts_dp_exchange_peer/event");
            tl_mark(s1, "landed", c);
            if ((rc = sh_gradient()) != TS_OK) return rc;
            // side 2: projection-backward over all views for MY shard of the piece
            if (ns > 0 && s2 != s1) TS_CHECK_CUDA(cudaStreamWaitEvent(s2, landed, 0), "ts_dp_exchange_peer/wait");
            if ((rc = shard_projection()) != TS_OK) return rc;
        }
    }
    // every shard's gradients have landed in my segment: final barrier on the side stream, then main joins
    if (s2 != s1) {
        cudaEvent_t e2 = pool_event(dev);
        TS_CHECK_CUDA(cudaEventRecord(e2, s2), "ts_dp_exchange_peer/event");
        TS_CHECK_CUDA(cudaStreamWaitEvent(s1, e2, 0), "ts_dp_exchange_peer/wait");
    }
    rc = ts_peer_barrier(world, rank, t_flags, ts::kBarrierSlots - 1, epoch, err, timeout_s, 3, side_stream);
    if (rc != TS_OK) return rc;
    tl_mark(s1, "all_landed", 0);
    if (s1 != sm) {
        cudaEvent_t e1 = pool_event(dev);
        TS_CHECK_CUDA(cudaEventRecord(e1, s1), "ts_dp_exchange_peer/event");
        TS_CHECK_CUDA(cudaStreamWaitEvent(sm, e1, 0), "ts_dp_exchange_peer/wait");
    }
    return TS_OK;
}
int ts_peer_flag_bytes(void) { return ts::kBarrierSlots * ts::kMaxPeers * 32 * (int)sizeof(uint32_t); }

}  // extern "C"
#endif  // !TS_HOST_EMU
