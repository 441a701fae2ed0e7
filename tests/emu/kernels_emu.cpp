// Compiles EVERY hot-path kernel of tinysplat_b200/csrc (project.cu, sh.cu, binning.cu, blend.cu,
// blend_group.cu) unchanged as host code on the fiber SIMT emulator (ts_emu.h) and exposes
//   * the blend kernels and the binning stages with the argument lists of their C-ABI entry points
//     (host pointers instead of device pointers; the launch logic mirrors those entry points), and
//   * emu_render_fused: the whole fused forward + backward of tinysplat_b200/fused.py,
//   * the SURVEY 8(f) widenings: fused Adam, fused SSIM forward/backward, brute-force knn_points.
// TEST INFRASTRUCTURE ONLY (TS_HOST_EMU is never defined in the product build).
#define TS_HOST_EMU 1
#include "../../tinysplat_b200/csrc/project.cu"
#include "../../tinysplat_b200/csrc/sh.cu"
#include "../../tinysplat_b200/csrc/binning.cu"
#include "../../tinysplat_b200/csrc/blend.cu"
#include "../../tinysplat_b200/csrc/blend_group.cu"
#include "../../tinysplat_b200/csrc/peer.cu"
#include "../../tinysplat_b200/csrc/adam.cu"
#include "../../tinysplat_b200/csrc/ssim.cu"
#include "../../tinysplat_b200/csrc/knn.cu"
#include "../../tinysplat_b200/csrc/loss.cu"

#include <vector>
#include <climits>

static int g_emu_key_cap = INT32_MAX;
static const int32_t* g_emu_tile_order = nullptr;   // launch order of the blend kernels (nullptr = raster order)

namespace {
template <int CH>
int run_fwd(int H, int W, int tx, int ty, const int32_t* off, const int32_t* ids, const float* recs,
            const float* bg, float* out_img, float* out_ch3, float* final_T, int32_t* n_contrib, int clamp) {
    const int32_t* order = g_emu_tile_order;
    return ts_emu::launch(dim3(tx * ty), ts::kBlendThreads, [=]() {
        ts::blend_fwd_kernel<CH>(H, W, tx, off, ids, (const float4*)recs, bg, out_img, out_ch3, final_T,
                                 n_contrib, clamp, g_emu_key_cap, order);
    });
}
template <int CH, int GCH>
int run_bwd(int grouped, int H, int W, int tx, int ty, const int32_t* off, const int32_t* ids, const float* recs,
            const float* bg, const float* final_T, const int32_t* n_contrib, const float* v_img,
            const float* v_ch3, int split, const float* v_alpha, float* grads) {
    const int32_t* order = g_emu_tile_order;
    if (grouped)
        return ts_emu::launch(dim3(tx * ty), ts::kGThreads, [=]() {
            ts::blend_bwd_group_kernel<CH, GCH>(H, W, tx, off, ids, (const float4*)recs, bg, final_T, n_contrib,
                                                v_img, v_ch3, split, v_alpha, (float4*)grads, order);
        });
    return ts_emu::launch(dim3(tx * ty), ts::kBlendThreads, [=]() {
        ts::blend_bwd_kernel<CH, GCH>(H, W, tx, off, ids, (const float4*)recs, bg, final_T, n_contrib,
                                      v_img, v_ch3, split, v_alpha, (float4*)grads, order);
    });
}
}  // namespace

extern "C" {

// the blend kernels launched after this call take their tiles in this order (nullptr: raster order)
void emu_set_tile_order(const int32_t* order) { g_emu_tile_order = order; }

int emu_bin_tile_order(int T, const int32_t* offsets, int32_t* order) {
    return ts_emu::launch(dim3(1), ts::kOrderThreads, [=]() { ts::bin_tile_order_kernel(T, offsets, order); });
}

int emu_blend_fwd(int CH, int H, int W, int tx, int ty, const int32_t* off, const int32_t* ids,
                  const float* recs, const float* bg, float* out_img, float* out_ch3, float* final_T,
                  int32_t* n_contrib, int clamp) {
    switch (CH) {
        case 1: return run_fwd<1>(H, W, tx, ty, off, ids, recs, bg, out_img, out_ch3, final_T, n_contrib, clamp);
        case 2: return run_fwd<2>(H, W, tx, ty, off, ids, recs, bg, out_img, out_ch3, final_T, n_contrib, clamp);
        case 3: return run_fwd<3>(H, W, tx, ty, off, ids, recs, bg, out_img, out_ch3, final_T, n_contrib, clamp);
        default: return run_fwd<4>(H, W, tx, ty, off, ids, recs, bg, out_img, out_ch3, final_T, n_contrib, clamp);
    }
}

// grouped: 0 = first-generation backward (blend.cu), 1 = grouped backward (blend_group.cu)
int emu_blend_bwd(int N, int CH, int H, int W, int tx, int ty, const int32_t* off, const int32_t* ids,
                  const float* recs, const float* bg, const float* final_T, const int32_t* n_contrib,
                  const float* v_img, const float* v_ch3, int split, const float* v_alpha, float* grads,
                  int grouped) {
    memset(grads, 0, sizeof(float) * ts::kGradFloats * (size_t)N);
    const int gch = (CH == 4 && split && !v_ch3) ? 3 : CH;
#define ARGS grouped, H, W, tx, ty, off, ids, recs, bg, final_T, n_contrib, v_img, v_ch3, split, v_alpha, grads
    switch (CH) {
        case 1: return run_bwd<1, 1>(ARGS);
        case 2: return run_bwd<2, 2>(ARGS);
        case 3: return run_bwd<3, 3>(ARGS);
        default: return gch == 3 ? run_bwd<4, 3>(ARGS) : run_bwd<4, 4>(ARGS);
    }
#undef ARGS
}

}  // extern "C"


extern "C" {

int emu_bin_counter_stride(void) { return ts::kCounterStride; }
int emu_bin_scan_work_ints(void) { return ts::kScanWorkInts; }
int emu_bin_smem_sort_cap(void) { return ts::kSmemSortCap; }

int emu_bin_count(int N, int CH, const float* xys, const int32_t* radii, const float* conics,
                  const float* opacity, const float* colors, int tx, int ty, int cull, int flags,
                  float* recs, int32_t* tile_counts) {
    memset(tile_counts, 0, sizeof(int32_t) * ts::kCounterStride * (size_t)tx * ty);
    if (N == 0) return 0;
    const int grid = (N + ts::kBinThreads - 1) / ts::kBinThreads;
#define RUN(C)                                                                                           \
    return ts_emu::launch(dim3(grid), ts::kBinThreads, [=]() {                                           \
        ts::bin_count_kernel<C>(N, (const float2*)xys, radii, conics, opacity, colors, tx, ty, cull, flags, \
                                (float4*)recs, tile_counts);                                             \
    })
    switch (CH) {
        case 1: RUN(1);
        case 2: RUN(2);
        case 3: RUN(3);
        default: RUN(4);
    }
#undef RUN
}

int emu_bin_scan(int T, int32_t* tile_counts, int32_t* offsets, int32_t* stats, int cap) {
    int chunk = (T + 1024 * ts::kScanMaxBlocks - 1) / (1024 * ts::kScanMaxBlocks);
    if (chunk < 1) chunk = 1;
    const int grid = (T + 1024 * chunk - 1) / (1024 * chunk);
    memset(stats, 0, sizeof(int32_t) * ts::kScanWorkInts);
    return ts_emu::launch(dim3(grid), 1024, [=]() { ts::bin_scan_kernel(T, chunk, tile_counts, offsets, stats, cap); });
}

int emu_bin_emit(int N, const float* depths, const int32_t* radii, const float* recs, int tx, int ty,
                 int cull, int32_t* cursors, uint64_t* keys) {
    if (N == 0) return 0;
    const int grid = (N + ts::kBinThreads - 1) / ts::kBinThreads;
    return ts_emu::launch(dim3(grid), ts::kBinThreads, [=]() {
        ts::bin_emit_kernel(N, depths, radii, (const float4*)recs, tx, ty, cull, cursors, keys, g_emu_key_cap);
    });
}

// capacity of the key / id buffers seen by emit, sort and blend-forward (tests of the overflow guards)
void emu_set_key_capacity(int cap) { g_emu_key_cap = cap > 0 ? cap : INT32_MAX; }

int emu_bin_reset_cursors(int T, const int32_t* offsets, int32_t* cursors) {
    return ts_emu::launch(dim3((T + 255) / 256), 256, [=]() { ts::bin_reset_cursors_kernel(T, offsets, cursors); });
}

int emu_bin_sort(int T, const int32_t* offsets, uint64_t* keys, int32_t* ids_sorted, int max_count,
                 int n_big, uint64_t* big_scratch, int32_t* big_counter) {
    if (max_count <= 0) return 0;
    const int cap = g_emu_key_cap;
    int rc = ts_emu::launch(dim3((T + 7) / 8), 256, [=]() { ts::bin_sort_warp_kernel(T, offsets, keys, ids_sorted, cap, n_big > 0 ? INT32_MAX : (max_count < ts::kSmemSortCap ? max_count : ts::kSmemSortCap)); });
    if (rc == 0 && max_count > ts::kWarpSortMax)
        rc = ts_emu::launch(dim3(T), 256, [=]() { ts::bin_sort_cta_kernel<256, 4, 8>(T, offsets, keys, ids_sorted, ts::kWarpSortMax, cap); });
    if (rc == 0 && max_count > 2048)
        rc = ts_emu::launch(dim3(T), 512, [=]() { ts::bin_sort_cta_kernel<512, 8, 16>(T, offsets, keys, ids_sorted, 2048, cap); });
    if (rc == 0 && max_count > 8192)
        rc = ts_emu::launch(dim3(T), 1024, [=]() { ts::bin_sort_cta_kernel<1024, 16, 16>(T, offsets, keys, ids_sorted, 8192, cap); });
    if (n_big > 0 && rc == 0) {
        int P = 2;
        while (P < max_count) P <<= 1;
        *big_counter = 0;
        rc = ts_emu::launch(dim3(T), 1024, [=]() {
            ts::bin_sort_big_kernel(T, offsets, keys, ids_sorted, ts::kSmemSortCap, P, big_scratch, big_counter, cap);
        });
    }
    return rc;
}

}  // extern "C"

// ---- the fused pipeline (tinysplat_b200/fused.py), stage by stage ------------------------------
namespace {

template <int DEG>
int sh_fwd_any(bool bulk, int N, int K, const float* means, const float* view, const float* dc, const float* rest,
               float* colors, int stride, const float* ch3, uint8_t* mask, int flags) {
    const int grid = (N + ts::kShThreads - 1) / ts::kShThreads;
    if (bulk)
        return ts_emu::launch(dim3(grid), ts::kShThreads, [=]() {
            ts::sh_fwd_bulk_kernel<DEG>(N, K, means, view, dc, rest, colors, stride, ch3, mask, flags);
        });
    const int sstride = ts::sh_stride(K);
    return ts_emu::launch(dim3(grid), ts::kShThreads, [=]() {
        ts::sh_fwd_kernel<DEG>(N, K, means, view, dc, rest, colors, stride, ch3, mask, flags, sstride);
    });
}

template <int DEG>
int sh_bwd_any(bool bulk, int N, int K, const float* means, const float* view, const float* v_colors, int stride,
               const uint8_t* mask, float* v_dc, float* v_rest, int flags) {
    const int grid = (N + ts::kShThreads - 1) / ts::kShThreads;
    if (bulk)
        return ts_emu::launch(dim3(grid), ts::kShThreads, [=]() {
            ts::sh_bwd_bulk_kernel<DEG>(N, K, means, view, v_colors, stride, mask, v_dc, v_rest, flags);
        });
    const int sstride = ts::sh_stride(K);
    return ts_emu::launch(dim3(grid), ts::kShThreads, [=]() {
        ts::sh_bwd_kernel<DEG>(N, K, means, view, v_colors, stride, mask, v_dc, v_rest, flags, sstride);
    });
}

template <int DEG>
int sh_bwd_views_any(int n_views, int N, int K, const float* means, const float* cams, const float* packed,
                     size_t view_stride, int row_stride, int col_off, float scale, float* v_dc, float* v_rest) {
    const int grid = (N + ts::kShThreads - 1) / ts::kShThreads;
    return ts_emu::launch(dim3(grid), ts::kShThreads, [=]() {
        ts::sh_bwd_views_kernel<DEG>(n_views, N, K, means, cams, packed, view_stride, row_stride, col_off, scale, v_dc,
                                     v_rest);
    });
}

template <int DEG>
int sh_bwd_views_rgb_any(int n_views, int N, int K, const float* means, const float* cams, const float* rgb,
                         size_t view_stride, float scale, float* v_dc, float* v_rest) {
    const int grid = (N + ts::kShThreads - 1) / ts::kShThreads;
    return ts_emu::launch(dim3(grid), ts::kShThreads, [=]() {
        ts::sh_bwd_views_rgb_kernel<DEG>(n_views, N, K, means, cams, rgb, view_stride, scale, v_dc, v_rest);
    });
}

template <int DEG>
int project_sh_bwd_any(int grid, int N, int K, const float* means, const float* log_scales, const float* quats,
                       const float* view, const float* fullproj, float fx, float fy, int W, int H, int flags,
                       const int32_t* radii, const float* grads, const float* logits, const uint8_t* mask,
                       float* v_means, float* v_scales, float* v_quats, float* v_logit, float* v_xys, float* v_dc,
                       float* v_rest) {
    return ts_emu::launch(dim3(grid), ts::kProjThreads, [=]() {
        ts::project_sh_bwd_kernel<DEG>(N, K, means, log_scales, 1.0f, (const float4*)quats, view, fullproj, fx, fy,
                                       W / 2.f, H / 2.f, H, W, flags, radii, (const float4*)grads, logits, mask, v_means,
                                       v_scales, (float4*)v_quats, v_logit, (float2*)v_xys, v_dc, v_rest);
    });
}

#define TS_EMU_BY_DEG(deg, fn, ...)             \
    switch (deg) {                              \
        case 0: return fn<0>(__VA_ARGS__);      \
        case 1: return fn<1>(__VA_ARGS__);      \
        case 2: return fn<2>(__VA_ARGS__);      \
        case 3: return fn<3>(__VA_ARGS__);      \
        default: return fn<4>(__VA_ARGS__);     \
    }

}  // namespace

extern "C" {

// Mirrors _RenderFused.forward + backward.  All pointers are host pointers; outputs are written by
// the kernels exactly as on the device.  v_rgb / v_depth may be NULL (no cotangent).
// bwd_mode bit 0: 0 = first-generation blend-backward, 1 = grouped; bit 1: K7 and K6 as two kernels instead of
// the fused ts_project_sh_bwd kernel (the product default).  Returns 0, or -1 on an emulator deadlock.
int emu_render_fused(int N, int K, int deg, int W, int H, const float* means, const float* log_scales,
                     const float* quats, const float* logits, const float* dc, const float* rest,
                     const float* view, const float* fullproj, float fx, float fy, const float* bg4, int cull,
                     int clamp_rgb, const float* v_rgb, const float* v_depth, int bwd_mode,
                     float* rgb, float* depth_img, float* final_T, float* xys, int32_t* radii, int64_t* stats_out,
                     float* v_means, float* v_scales, float* v_quats, float* v_logit, float* v_dc, float* v_rest,
                     float* v_xys) {
    const int tx = (W + 15) / 16, ty = (H + 15) / 16, T = tx * ty;
    const int pflags = TS_PROJ_LOG_SCALES | TS_PROJ_RAW_QUATS;
    const int sflags = TS_SH_DIRS_FROM_MEANS | TS_SH_OFFSET_CLAMP;
    std::vector<float> depths(N + 1), recs((size_t)N * 12 + 4, 0.f);
    std::vector<int32_t> counts((size_t)T * ts::kCounterStride, 0), offsets(T + 1, 0), stats(ts::kScanWorkInts, 0);
    std::vector<uint8_t> mask(N + 1, 0);
    float* recs_p = (float*)(((uintptr_t)recs.data() + 15) & ~(uintptr_t)15);
    int rc = 0;
    // K1 + pack + count
    if (N > 0) {
        const int grid = (N + ts::kProjThreads - 1) / ts::kProjThreads;
        rc = ts_emu::launch(dim3(grid), ts::kProjThreads, [&]() {
            ts::project_fwd_kernel(N, means, log_scales, 1.0f, (const float4*)quats, view, fullproj, fx, fy, W / 2.f,
                                   H / 2.f, H, W, tx, ty, 0.01f, pflags | TS_PROJ_OPACITY_LOGIT, (float2*)xys,
                                   depths.data(), radii, nullptr, nullptr, nullptr, logits, cull, (float4*)recs_p,
                                   counts.data());
        });
        if (rc) return rc;
        // K2: colours straight into the third float4 of the record, depth as channel 3
        const bool bulk = K == (deg + 1) * (deg + 1) && K > 1 && ((K - 1) * 3 * 4 * ts::kShThreads) % 16 == 0;
        auto run = [&]() -> int {
            TS_EMU_BY_DEG(deg, sh_fwd_any, bulk, N, K, means, view, dc, rest, recs_p + 8, 12, depths.data(), mask.data(), sflags);
        };
        rc = run();
        if (rc) return rc;
    }
    // K3
    rc = emu_bin_scan(T, counts.data(), offsets.data(), stats.data(), ts::kSmemSortCap);
    if (rc) return rc;
    const int M = stats[0], max_count = stats[1], n_big = stats[2];
    stats_out[0] = M; stats_out[1] = max_count; stats_out[2] = n_big;
    std::vector<uint64_t> keys(M + 1, 0);
    std::vector<int32_t> ids(M + 1, 0);
    if (M > 0) {
        rc = emu_bin_emit(N, depths.data(), radii, recs_p, tx, ty, cull, counts.data(), keys.data());
        if (rc) return rc;
        std::vector<uint64_t> scratch;
        int32_t counter = 0;
        if (n_big > 0) {
            int P = 2;
            while (P < max_count) P <<= 1;
            scratch.resize((size_t)n_big * P);
        }
        rc = emu_bin_sort(T, offsets.data(), keys.data(), ids.data(), max_count, n_big, scratch.data(), &counter);
        if (rc) return rc;
    }
    // launch order of the blend kernels, as tinysplat_b200/binning.py computes it
    std::vector<int32_t> order(T, 0);
    rc = emu_bin_tile_order(T, offsets.data(), order.data());
    if (rc) return rc;
    struct OrderScope {
        explicit OrderScope(const int32_t* o) { g_emu_tile_order = o; }
        ~OrderScope() { g_emu_tile_order = nullptr; }
    } order_scope(order.data());
    // K4
    std::vector<int32_t> ncon((size_t)W * H, 0);
    rc = emu_blend_fwd(4, H, W, tx, ty, offsets.data(), ids.data(), recs_p, bg4, rgb, depth_img, final_T, ncon.data(),
                       clamp_rgb);
    if (rc || (!v_rgb && !v_depth)) return rc;
    // K5
    std::vector<float> grads((size_t)N * 12 + 4, 0.f);
    float* grads_p = (float*)(((uintptr_t)grads.data() + 15) & ~(uintptr_t)15);
    if (N == 0) return 0;
    rc = emu_blend_bwd(N, 4, H, W, tx, ty, offsets.data(), ids.data(), recs_p, bg4, final_T, ncon.data(), v_rgb, v_depth,
                       1, nullptr, grads_p, bwd_mode & 1);
    if (rc) return rc;
    // K6 + K7 as one kernel (fused.py FUSED_TAIL)
    if (!(bwd_mode & 2)) {
        const int grid = (N + ts::kProjThreads - 1) / ts::kProjThreads;
        auto run = [&]() -> int {
            TS_EMU_BY_DEG(deg, project_sh_bwd_any, grid, N, K, means, log_scales, quats, view, fullproj, fx, fy, W, H,
                          pflags | TS_PROJ_DEPTH_CH3, radii, grads_p, logits, mask.data(), v_means, v_scales, v_quats,
                          v_logit, v_xys, v_dc, v_rest);
        };
        return run();
    }
    // K7, K6
    {
        const bool bulk = K > 1 && ((K - 1) * 3 * 4 * ts::kShThreads) % 16 == 0;
        auto run = [&]() -> int {
            TS_EMU_BY_DEG(deg, sh_bwd_any, bulk, N, K, means, view, grads_p + 8, 12, mask.data(), v_dc, v_rest, sflags);
        };
        rc = run();
        if (rc) return rc;
    }
    const int grid = (N + ts::kProjThreads - 1) / ts::kProjThreads;
    return ts_emu::launch(dim3(grid), ts::kProjThreads, [&]() {
        ts::project_bwd_kernel(N, means, log_scales, 1.0f, (const float4*)quats, view, fullproj, fx, fy, W / 2.f, H / 2.f,
                               H, W, pflags | TS_PROJ_DEPTH_CH3, radii, nullptr, nullptr, nullptr,
                               (const float4*)grads_p, logits, v_means, v_scales, (float4*)v_quats, v_logit,
                               (float2*)v_xys);
    });
}

// The shard backward of the packed-row gradient exchange (ts_project_bwd_views + ts_sh_bwd_views).
int emu_shard_bwd_views(int n_views, int N, int K, int deg, int W, int H, const float* means,
                        const float* log_scales, const float* quats, const float* logits, const float* cams,
                        const float* packed, int64_t view_stride, float scale, float* v_means, float* v_scales,
                        float* v_quats, float* v_logit, float* v_dc, float* v_rest) {
    if (N == 0) return 0;
    const int flags = TS_PROJ_LOG_SCALES | TS_PROJ_RAW_QUATS | TS_PROJ_DEPTH_CH3;
    const int grid = (N + ts::kProjThreads - 1) / ts::kProjThreads;
    ts::PeerPtrs dm{}, ds{}, dq{}, dl{};
    dm.p[0] = v_means; ds.p[0] = v_scales; dq.p[0] = v_quats; dl.p[0] = v_logit;
    int rc = ts_emu::launch(dim3(grid), ts::kProjThreads, [&]() {
        ts::project_bwd_views_kernel<false>(n_views, N, means, log_scales, 1.0f, (const float4*)quats, cams, H, W, flags,
                                            (const float4*)packed, (size_t)(view_stride / 4), logits, scale, 1, 0, dm, ds,
                                            dq, dl);
    });
    if (rc) return rc;
    auto run = [&]() -> int {
        TS_EMU_BY_DEG(deg, sh_bwd_views_any, n_views, N, K, means, cams, packed, (size_t)view_stride, 12, 8, scale, v_dc,
                      v_rest);
    };
    return run();
}

// The peer-memory gradient exchange (peer.cu) with `world` ranks simulated in one address space:
// every rank's ts_dp_push writes into every other rank's buffers, then every rank runs the colour
// part (SH-backward over all views, all Gaussians) and the shard part (projection-backward over all
// views for its shard, stored into EVERY rank's gradient arrays).  Inputs with a leading [world]
// dimension are per rank/view: packed rows [world][N][12], radii [world][N], mask [world][N],
// recs [world][N][12], cams [world][32].  Outputs [world][...]: what each rank ends up holding.
int emu_peer_exchange(int world, int N, int Ns, int K, int deg, int W, int H, const float* means,
                      const float* log_scales, const float* quats, const float* logits, const float* packed_all,
                      const int32_t* radii_all, const uint8_t* mask_all, const float* recs_all,
                      const float* cams_all, float scale, float* v_means_all, float* v_scales_all,
                      float* v_quats_all, float* v_logit_all, float* v_dc_all, float* v_rest_all, float* v_xys_all) {
    if (world < 1 || world > ts::kMaxPeers || Ns % 4 != 0 || (int64_t)Ns * world < N) return -2;
    const int Npad = world * Ns;
    const int R = (K - 1) * 3;
    std::vector<std::vector<float>> geo(world), rgb(world), cams(world);
    ts::PeerPtrs pgeo{}, prgb{}, pcams{};
    auto align16 = [](std::vector<float>& v) { return (float*)(((uintptr_t)v.data() + 15) & ~(uintptr_t)15); };
    for (int r = 0; r < world; ++r) {
        geo[r].assign((size_t)world * Ns * 8 + 4, -777.f);      // poisoned: unwritten rows must never be read
        rgb[r].assign((size_t)world * Npad * 3 + 4, -777.f);
        cams[r].assign((size_t)world * 32 + 4, -777.f);
        pgeo.p[r] = align16(geo[r]); prgb.p[r] = align16(rgb[r]); pcams.p[r] = align16(cams[r]);
    }
    // persistent grid: fewer CTAs than 256-row blocks, so that every CTA iterates and reuses both of its buffers
    const int pblocks = (N + ts::kPushRows - 1) / ts::kPushRows;
    const int pgrid = pblocks > 3 ? (pblocks + 2) / 3 : 1;
    // odd ranks push geometry and colour rows in two passes (what = 1, then 2: the split exchange), even ranks in one
    for (int r = 0; r < world; ++r) {
        for (int pass = 0; pass < ((r & 1) ? 2 : 1); ++pass) {
            const int what = (r & 1) ? pass + 1 : 3;
            int rc = ts_emu::launch(dim3(pgrid), ts::kPushThreads, [&]() {
                ts::dp_push_kernel(N, Ns, Npad, world, r, radii_all + (size_t)r * N, mask_all + (size_t)r * N,
                                   (const float4*)(recs_all + (size_t)r * N * 12), (const float4*)(packed_all + (size_t)r * N * 12),
                                   cams_all + (size_t)r * 32, pgeo, prgb, pcams, (float2*)(v_xys_all + (size_t)r * N * 2), what);
            });
            if (rc) return rc;
        }
    }
    const int flags = TS_PROJ_LOG_SCALES | TS_PROJ_RAW_QUATS;
    for (int r = 0; r < world; ++r) {
        // colour part: every rank, all Gaussians
        if (N > 0) {
            float* vdc = v_dc_all + (size_t)r * N * 3;
            float* vrest = v_rest_all + (size_t)r * N * R;
            const float* rows = (const float*)prgb.p[r];
            const float* cam = (const float*)pcams.p[r];
            auto run = [&]() -> int {
                TS_EMU_BY_DEG(deg, sh_bwd_views_rgb_any, world, N, K, means, cam, rows, (size_t)Npad * 3, scale, vdc, vrest);
            };
            int rc = run();
            if (rc) return rc;
        }
        // shard part: rank r owns rows [s0, s0 + ns)
        const int s0 = r * Ns;
        const int ns = std::max(0, std::min(N, s0 + Ns) - s0);
        if (ns == 0) continue;
        ts::PeerPtrs dm{}, ds{}, dq{}, dl{};
        for (int d = 0; d < world; ++d) {
            dm.p[d] = v_means_all + ((size_t)d * N + s0) * 3;
            ds.p[d] = v_scales_all + ((size_t)d * N + s0) * 3;
            dq.p[d] = v_quats_all + ((size_t)d * N + s0) * 4;
            dl.p[d] = v_logit_all + ((size_t)d * N + s0);
        }
        const int grid = (ns + ts::kProjThreads - 1) / ts::kProjThreads;
        int rc = ts_emu::launch(dim3(grid), ts::kProjThreads, [&]() {
            ts::project_bwd_views_kernel<true>(world, ns, means + (size_t)s0 * 3, log_scales + (size_t)s0 * 3, 1.0f,
                                               (const float4*)(quats + (size_t)s0 * 4), (const float*)pcams.p[r], H, W,
                                               flags, (const float4*)pgeo.p[r], (size_t)Ns * 2, logits + s0, scale, world,
                                               (r + 1) % world, dm, ds, dq, dl);
        });
        if (rc) return rc;
    }
    return 0;
}

// ts_project_fwd (gsplat contract: no packing / counting) on host pointers.
int emu_project_fwd(int N, const float* means, const float* scales, const float* quats, const float* view,
                    const float* fullproj, float fx, float fy, int W, int H, int flags, float* xys, float* depths,
                    int32_t* radii, float* conics, int32_t* ntiles, float* cov3d) {
    if (N == 0) return 0;
    const int tx = (W + 15) / 16, ty = (H + 15) / 16;
    const int grid = (N + ts::kProjThreads - 1) / ts::kProjThreads;
    return ts_emu::launch(dim3(grid), ts::kProjThreads, [=]() {
        ts::project_fwd_kernel(N, means, scales, 1.0f, (const float4*)quats, view, fullproj, fx, fy, W / 2.f, H / 2.f, H,
                               W, tx, ty, 0.01f, flags, (float2*)xys, depths, radii, conics, ntiles, cov3d, nullptr, 0,
                               nullptr, nullptr);
    });
}

// ts_project_bwd on host pointers (explicit cotangents and/or packed rows, like the C-ABI entry).
int emu_project_bwd(int N, const float* means, const float* scales, const float* quats, const float* view,
                    const float* fullproj, float fx, float fy, int W, int H, int flags, const int32_t* radii,
                    const float* v_xys, const float* v_depths, const float* v_conics, const float* packed,
                    const float* logits, float* v_means, float* v_scales, float* v_quats, float* v_logit,
                    float* v_xys_out) {
    if (N == 0) return 0;
    const int grid = (N + ts::kProjThreads - 1) / ts::kProjThreads;
    return ts_emu::launch(dim3(grid), ts::kProjThreads, [=]() {
        ts::project_bwd_kernel(N, means, scales, 1.0f, (const float4*)quats, view, fullproj, fx, fy, W / 2.f, H / 2.f, H,
                               W, flags, radii, (const float2*)v_xys, v_depths, v_conics, (const float4*)packed, logits,
                               v_means, v_scales, (float4*)v_quats, v_logit, (float2*)v_xys_out);
    });
}

// ts_dp_prepare on host pointers.
int emu_dp_prepare(int N, const int32_t* radii, const uint8_t* mask, const float* recs, float* grads, float* v_xys) {
    if (N == 0) return 0;
    return ts_emu::launch(dim3((N + 255) / 256), 256, [=]() {
        ts::dp_prepare_kernel(N, radii, mask, (const float4*)recs, (float4*)grads, (float2*)v_xys);
    });
}

}  // extern "C"

// ---- Adam, SSIM, knn_points -----------------------------------------------------
extern "C" {

int emu_adam_step(int num_tensors, float* const* params, const float* const* grads, float* const* exp_avgs,
                  float* const* exp_avg_sqs, const int64_t* numels, const float* lrs, const int64_t* steps,
                  double beta1, double beta2, double eps) {
    ts::AdamTensors t;
    int blocks = 0;
    int rc = ts::adam_build(num_tensors, params, grads, exp_avgs, exp_avg_sqs, numels, lrs, steps, beta1, beta2, t,
                            blocks);
    if (rc != 0 || blocks == 0) return rc;
    const float omb1 = (float)(1.0 - beta1), b2 = (float)beta2, omb2 = (float)(1.0 - beta2), e = (float)eps;
    return ts_emu::launch(dim3(blocks), ts::kAdamThreads, [=]() { ts::adam_multi_kernel(t, omb1, b2, omb2, e); });
}

int emu_ssim_fwd(int B, int C, int H, int W, const float* X, const int64_t* xs, const float* Y, const int64_t* ys,
                 const float* win11, float C1, float C2, float* ssim_sum, float* dmu, float* de11, float* de12) {
    memset(ssim_sum, 0, sizeof(float) * (size_t)B * C);
    ts::SsimWin win;
    for (int k = 0; k < ts::kWin; ++k) win.w[k] = win11[k];
    ts::Strides sx{xs[0], xs[1], xs[2], xs[3]}, sy{ys[0], ys[1], ys[2], ys[3]};
    const int Ho = H - ts::kHalo, Wo = W - ts::kHalo;
    dim3 grid((Wo + ts::kST - 1) / ts::kST, (Ho + ts::kST - 1) / ts::kST, B * C), block(ts::kST, ts::kST);
    return ts_emu::launch(grid, block, [=]() {
        ts::ssim_fwd_kernel(C, H, W, X, Y, sx, sy, win, C1, C2, ssim_sum, dmu, de11, de12);
    });
}

int emu_ssim_bwd(int B, int C, int H, int W, const float* X, const int64_t* xs, const float* Y, const int64_t* ys,
                 const float* win11, const float* dmu, const float* de11, const float* de12, const float* v_pc,
                 float* v_X) {
    ts::SsimWin win;
    for (int k = 0; k < ts::kWin; ++k) win.w[k] = win11[k];
    ts::Strides sx{xs[0], xs[1], xs[2], xs[3]}, sy{ys[0], ys[1], ys[2], ys[3]};
    dim3 grid((W + ts::kST - 1) / ts::kST, (H + ts::kST - 1) / ts::kST, B * C), block(ts::kST, ts::kST);
    return ts_emu::launch(grid, block, [=]() {
        ts::ssim_bwd_kernel(C, H, W, X, Y, sx, sy, win, dmu, de11, de12, v_pc, v_X);
    });
}

// fused L1 loss: `blocks` CTAs grid-stride over the image, the last one to finish sums the partials
int emu_l1_loss(long long n, const float* img, const void* target, int is_u8, float grad_scale, float loss_scale,
                float* grad, float* work, float* loss, int blocks) {
    unsigned int* counter = reinterpret_cast<unsigned int*>(work + ts::kLossMaxBlocks);
    if (is_u8)
        return ts_emu::launch(dim3(blocks), ts::kLossThreads, [=]() {
            ts::l1_loss_kernel<uint8_t>((size_t)n, img, (const uint8_t*)target, grad_scale, loss_scale, grad, work, counter, loss);
        });
    return ts_emu::launch(dim3(blocks), ts::kLossThreads, [=]() {
        ts::l1_loss_kernel<float>((size_t)n, img, (const float*)target, grad_scale, loss_scale, grad, work, counter, loss);
    });
}
int emu_l1_loss_work_floats() { return ts::kLossMaxBlocks + 4; }

int emu_knn_points(int P1, int P2, int K, const float* q, const float* ref, float* dists, int64_t* idx) {
    if (P1 == 0) return 0;
    const int grid = (P1 + ts::kKnnThreads - 1) / ts::kKnnThreads;
#define RUN(KK) return ts_emu::launch(dim3(grid), ts::kKnnThreads, [=]() { ts::knn_kernel<KK>(P1, P2, q, ref, dists, idx); })
    switch (K) {
        case 1: RUN(1);
        case 2: RUN(2);
        case 4: RUN(4);
        case 8: RUN(8);
        case 16: RUN(16);
        case 32: RUN(32);
        default: return -1;
    }
#undef RUN
}

}  // extern "C"
