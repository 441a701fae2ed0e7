/* tinysplat_b200 — C ABI of the B200 (sm_100a) Gaussian-splatting hot path.
 *
 * This is the drop-in boundary underneath the five `gsplat` symbols tinysplat imports
 * [REF tinysplat/splatting/rasterize.py:3-4; tinysplat/splatting/model_gaussian.py:14].
 * Every entry point is `extern "C"`, takes plain device pointers + sizes + a CUDA stream,
 * returns 0 on success or a negative ts_status, allocates nothing persistent and never
 * throws.  The caller owns every buffer.  All pointers are DEVICE pointers unless the
 * parameter name ends in `_host`.  All float arrays are fp32, row-major, contiguous.
 * Pointers marked [16B] must be 16-byte aligned (any fresh torch allocation is).
 *
 * The reference-side binding is a ctypes stub: see INTEGRATION.md and
 * tinysplat_b200/_lib.py.
 */
#ifndef TINYSPLAT_B200_H
#define TINYSPLAT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define TS_API __declspec(dllexport)
#else
#define TS_API __attribute__((visibility("default")))
#endif

typedef void* ts_stream_t; /* cudaStream_t */

enum ts_status {
    TS_OK = 0,
    TS_ERR_INVALID = -1,   /* bad argument (size, channel count, degree, null pointer) */
    TS_ERR_ALIGN = -2,     /* a [16B] pointer is misaligned */
    TS_ERR_CUDA = -3,      /* a CUDA runtime call or launch failed; see ts_last_error() */
    TS_ERR_CAPACITY = -4   /* a caller-provided workspace is too small */
};

/* Flags of the fused pipeline: the activations the reference adapter applies in torch around
 * the gsplat calls [REF rasterize.py:39,72-73,77-79,86] folded into the kernels. */
enum ts_flags {
    TS_PROJ_LOG_SCALES = 1,   /* scales are log-scales: exp() inside, Jacobian in backward */
    TS_PROJ_RAW_QUATS = 2,    /* quats are unnormalised: q/|q| inside, Jacobian in backward */
    TS_PROJ_DEPTH_CH3 = 4,    /* backward: depth cotangent = colour channel 3 of packed grads */
    TS_PROJ_OPACITY_LOGIT = 8,/* forward pack+count: `opacity` holds logits: sigmoid() inside */
    TS_SH_DIRS_FROM_MEANS = 1,/* `dirs` holds means3d; dir = mean - viewmat[:3,3] */
    TS_SH_OFFSET_CLAMP = 2,   /* colour = max(sh + 0.5, 0); mask of passing channels saved */
    TS_BIN_OPACITY_LOGIT = 1, /* `opacity` holds logits: sigmoid() inside */
    TS_BIN_PACK_ONLY = 2,     /* ts_bin_count: only pack the records (the caller reuses tile lists it already has) */
    TS_BLEND_GRADS_ZEROED = 2 /* ts_blend_bwd, or-ed into split_ch3: the caller has already zeroed `grads` (e.g. on an
                                 idle stream during the forward pass); the callee then skips its memset */
};

/* Library version (major*10000 + minor*100 + patch) and last CUDA error text. */
TS_API int ts_version(void);
TS_API const char* ts_last_error(void);
/* Number of floats in one packed per-Gaussian raster record / gradient record. */
TS_API int ts_rec_floats(void);
TS_API int ts_grad_floats(void);
/* Number of kernels this library has launched since load (for bench.py's gpu_launches). */
TS_API int64_t ts_launch_count(void);

/* ---- K1: EWA projection forward ----------------------------------------------------
 * Replaces gsplat.project_gaussians  [REF rasterize.py:32, args marshalled at :64-73].
 * viewmat: 3x4 (first three rows of the 4x4 world->camera matrix), projmat: 4x4 full
 * projection (proj @ view).  Outputs: xys[N,2], depths[N], radii[N] (int32),
 * conics[N,3], num_tiles_hit[N] (int32), cov3d[N,6].  Culled Gaussians get zeros.
 * flags: TS_PROJ_LOG_SCALES | TS_PROJ_RAW_QUATS (0 = the gsplat contract).
 * Fused pipeline: with recs != NULL the kernel also packs the geometry half of the raster
 * record (recs[N, ts_rec_floats()], floats 0..7) and counts tile intersections into
 * tile_counts[T] (zeroed by the callee) exactly like ts_bin_count, from `opacity[N]`
 * (logits with TS_PROJ_OPACITY_LOGIT) and cull_mode; conics / num_tiles_hit / cov3d may then
 * be NULL (not written). */
TS_API int ts_project_fwd(int N,
                          const float* means3d /*[16B]*/, const float* scales /*[16B]*/,
                          float glob_scale, const float* quats /*[16B]*/,
                          const float* viewmat, const float* projmat,
                          float fx, float fy, float cx, float cy,
                          int img_height, int img_width, int tiles_x, int tiles_y,
                          float clip_thresh, int flags,
                          float* xys /*[16B]*/, float* depths, int32_t* radii,
                          float* conics /*[16B] or NULL*/, int32_t* num_tiles_hit /*or NULL*/,
                          float* cov3d /*[16B] or NULL*/,
                          const float* opacity /*or NULL*/, int cull_mode,
                          float* recs /*[16B] or NULL*/, int32_t* tile_counts /*or NULL*/,
                          ts_stream_t stream);

/* ---- K6: EWA projection backward ---------------------------------------------------
 * Backward of ts_project_fwd: consumes v_xys[N,2], v_depths[N], v_conics[N,3] (each may be
 * NULL = zero; the depth cotangent is live because the reference rasterises depth as a
 * colour [REF rasterize.py:48-51]) and produces v_means3d[N,3], v_scales[N,3], v_quats[N,4].
 * Fused pipeline: packed_grads[N, ts_grad_floats()] (blend-backward's output) is consumed
 * directly and ADDED to the explicit cotangents; with opacity_logits/v_opacity_logits the
 * sigmoid Jacobian is applied here too; v_xys_out[N,2] (optional) receives d loss / d xy, the
 * densification statistic's input [REF model_gaussian.py:130-132]. */
TS_API int ts_project_bwd(int N,
                          const float* means3d /*[16B]*/, const float* scales /*[16B]*/,
                          float glob_scale, const float* quats /*[16B]*/,
                          const float* viewmat, const float* projmat,
                          float fx, float fy, float cx, float cy,
                          int img_height, int img_width, int flags,
                          const int32_t* radii,
                          const float* v_xys /*or NULL*/, const float* v_depths /*or NULL*/,
                          const float* v_conics /*[16B] or NULL*/,
                          const float* packed_grads /*[16B] or NULL*/,
                          const float* opacity_logits /*or NULL*/,
                          float* v_means3d /*[16B]*/, float* v_scales /*[16B]*/,
                          float* v_quats /*[16B]*/, float* v_opacity_logits /*or NULL*/,
                          float* v_xys_out /*or NULL*/, ts_stream_t stream);

/* ---- K2/K7: spherical harmonics ----------------------------------------------------
 * Replaces gsplat.sh.spherical_harmonics  [REF rasterize.py:38, args at :75-81].
 * coeffs is [N,K,3]; the first (degree+1)^2 bases are used.
 * If coeffs_rest is non-null, coeffs holds only the DC band [N,1,3] and coeffs_rest the
 * remaining [N,K-1,3] (the reference stores them as two Parameters
 * [REF model_gaussian.py:86-87] and concatenates every step [REF rasterize.py:80]).
 * colors: out_stride floats per Gaussian (3 = plain [N,3]; 12 with colors = recs+8 writes
 * straight into the packed raster record, channel 3 <- ch3[N] if given).
 * flags: TS_SH_DIRS_FROM_MEANS (needs viewmat) | TS_SH_OFFSET_CLAMP (clamp_mask[N] out). */
TS_API int ts_sh_fwd(int N, int degree, int K, const float* dirs /*[16B]*/,
                     const float* viewmat /*or NULL*/,
                     const float* coeffs /*[16B]*/, const float* coeffs_rest /*[16B] or NULL*/,
                     float* colors /*[16B]*/, int out_stride, const float* ch3 /*or NULL*/,
                     uint8_t* clamp_mask /*or NULL*/, int flags, ts_stream_t stream);
TS_API int ts_sh_bwd(int N, int degree, int K, const float* dirs /*[16B]*/,
                     const float* viewmat /*or NULL*/,
                     const float* v_colors /*[16B]*/, int v_stride,
                     const uint8_t* clamp_mask /*or NULL*/,
                     float* v_coeffs /*[16B]*/, float* v_coeffs_rest /*[16B] or NULL*/,
                     int flags, ts_stream_t stream);

/* ---- K3: tile binning + per-tile depth sort ------------------------------------------
 * Inside gsplat.rasterize_gaussians  [REF rasterize.py:44,50: no bins are passed in, so
 * binning happens behind the call].  Four steps; the caller reads stats_host between
 * ts_bin_scan and ts_bin_emit to size `keys` / `ids_sorted`.
 *
 * ts_bin_count : packs one raster record per Gaussian (recs[N, ts_rec_floats()]; the colour
 *                float4 is skipped when colors == NULL, ts_sh_fwd then writes it) and
 *                counts, per tile, the Gaussians whose footprint can reach a pixel of it
 *                (tile_counts[T * ts_bin_counter_stride()], zeroed by the callee).  CH = colour
 *                channels (1..4).
 * ts_bin_scan  : exclusive scan -> tile_offsets[T+1]; stats[0..3] = {total M, max per-tile
 *                count, number of tiles whose count exceeds smem_sort_cap, 0}.  `stats` must
 *                hold ts_bin_scan_work_ints() int32 (the tail is scan workspace; zeroed by
 *                the callee).
 * ts_bin_emit  : writes keys[M] = depth_bits<<32 | gaussian_id grouped by tile
 *                (cursors = the tile_counts buffer after ts_bin_scan).
 * ts_bin_sort  : sorts every tile's keys and writes ids_sorted[M] (gaussian ids, front to back,
 *                ties by gaussian id).  max_count / n_big_tiles select the kernels to launch (an
 *                upper bound is fine; a tile longer than max_count is left unsorted).  big_scratch
 *                must hold n_big_tiles * next_pow2(max_count) uint64 when n_big_tiles > 0.
 * `capacity` (ts_bin_emit, ts_bin_sort, ts_blend_fwd): number of entries the caller's keys /
 *   ids_sorted buffers hold, or 0 = exactly M.  A caller that sizes the buffers BEFORE it has read
 *   M back (from an earlier step, with headroom) passes the capacity: entries past it are never
 *   written or read; if M turns out larger the results are void and the caller must enlarge the
 *   buffers, restore the cursors (ts_bin_reset_cursors) and run emit / sort / blend again. */
TS_API int ts_bin_count(int N, int CH, const float* xys, const float* depths,
                        const int32_t* radii, const float* conics, const float* opacity,
                        const float* colors, int img_height, int img_width,
                        int tiles_x, int tiles_y, int cull_mode, int flags,
                        float* recs /*[16B]*/, int32_t* tile_counts, ts_stream_t stream);
TS_API int ts_bin_scan(int num_tiles, int32_t* tile_counts, int32_t* tile_offsets,
                       int32_t* stats, int smem_sort_cap, ts_stream_t stream);
TS_API int ts_bin_emit(int N, const float* depths, const int32_t* radii,
                       const float* recs /*[16B]*/, int tiles_x, int tiles_y, int cull_mode,
                       const int32_t* tile_offsets, int32_t* cursors, uint64_t* keys,
                       int capacity, ts_stream_t stream);
TS_API int ts_bin_reset_cursors(int num_tiles, const int32_t* tile_offsets, int32_t* cursors,
                                ts_stream_t stream);
/* tile_order[num_tiles]: a permutation of the tiles by descending list length (256 length classes),
 * the launch order of the blend kernels (see ts_blend_fwd). */
TS_API int ts_bin_tile_order(int num_tiles, const int32_t* tile_offsets, int32_t* tile_order,
                             ts_stream_t stream);
TS_API int ts_bin_sort(int num_tiles, const int32_t* tile_offsets, uint64_t* keys,
                       int32_t* ids_sorted, int max_count, int n_big_tiles,
                       uint64_t* big_scratch, int32_t* big_counter, int capacity,
                       ts_stream_t stream);
/* Per-tile counters are strided: tile t's counter is tile_counts[t * ts_bin_counter_stride()]
 * (one per 128-byte line: L2 serialises atomics per line).  ts_bin_scan rewrites each counter
 * in place with the tile's exclusive offset, so the same buffer is ts_bin_emit's `cursors`. */
TS_API int ts_bin_counter_stride(void);
TS_API int ts_bin_scan_work_ints(void);
/* Largest per-tile list ts_bin_sort sorts in shared memory. */
TS_API int ts_bin_smem_sort_cap(void);

/* ---- K4/K5: alpha-blend forward / backward -----------------------------------------
 * ts_blend_fwd: per-pixel front-to-back compositing of each tile's sorted list.
 *   out_img[H,W,CH], final_T[H,W], n_contrib[H,W] (int32: position after the last
 *   contributing list entry; what backward replays from).  With CH == 4 and out_ch3 != NULL
 *   the output is split: out_img[H,W,3] + out_ch3[H,W] (RGB + depth of the fused pass);
 *   backward mirrors it with split_ch3 = 1 (v_out_img[H,W,3] / v_out_ch3[H,W], NULL = 0).
 *   clamp_max1 (split mode only) folds the adapter's clamp(rgb, max=1) [REF rasterize.py:45]
 *   in; the clamped-channel mask rides in bits 28..30 of n_contrib and zeroes those
 *   cotangents in backward.
 * ts_blend_bwd: replays back to front; accumulates per-Gaussian packed gradients
 *   grads[N, ts_grad_floats()] (zeroed by the callee unless split_ch3 carries TS_BLEND_GRADS_ZEROED).
 *   v_out_alpha may be NULL.
 * ts_blend_unpack_grads: packed -> v_xys[N,2], v_conics[N,3], v_colors[N,CH], v_opacity[N].
 * tile_order (both blend calls): ts_bin_tile_order's permutation — CTA i works on tile
 *   tile_order[i], longest lists first, so the grid's tail is made of cheap tiles; NULL = raster
 *   order.  The results do not depend on it. */
TS_API int ts_blend_fwd(int CH, int img_height, int img_width, int tiles_x, int tiles_y,
                        const int32_t* tile_offsets, const int32_t* ids_sorted,
                        const float* recs /*[16B]*/, const float* background,
                        float* out_img, float* out_ch3 /*or NULL*/, float* final_T,
                        int32_t* n_contrib, int clamp_max1, int capacity,
                        const int32_t* tile_order /*or NULL*/, ts_stream_t stream);
TS_API int ts_blend_bwd(int N, int CH, int img_height, int img_width, int tiles_x, int tiles_y,
                        const int32_t* tile_offsets, const int32_t* ids_sorted,
                        const float* recs /*[16B]*/, const float* background,
                        const float* final_T, const int32_t* n_contrib,
                        const float* v_out_img, const float* v_out_ch3 /*or NULL*/,
                        int split_ch3, const float* v_out_alpha /*or NULL*/,
                        float* grads /*[16B]*/, const int32_t* tile_order /*or NULL*/,
                        ts_stream_t stream);
TS_API int ts_blend_unpack_grads(int N, int CH, const int32_t* radii, const float* conics,
                                 const float* grads /*[16B]*/, float* v_xys, float* v_conics,
                                 float* v_colors, float* v_opacity, ts_stream_t stream);
/* Two backward kernels sit behind ts_blend_bwd with identical semantics: mode 1 (default) =
 * grouped: one 8-lane group per 8x4 sub-block, four rows per lane, exact per-row culling, group
 * totals sent to global memory as vector reds (blend_group.cu); mode 0 = first generation: one
 * warp per sub-block, bounding-box culling, shared-memory group reduction (blend.cu).  The default
 * comes from the environment variable TS_BLEND_MODE ("warp" | "group") or the built-in default;
 * ts_set_blend_mode(-1) returns to it.  Process-wide, not thread-safe against concurrent
 * launches. */
TS_API int ts_set_blend_mode(int mode);
TS_API int ts_get_blend_mode(void);
/* Test hook (host code, no GPU): the exact row mask the grouped backward computes for one packed
 * record (q0 = {x, y, hx, hy}, q1 = {A, B, C, opacity}, see ts_rec_floats) against tile
 * (tile_x, tile_y): bit (2*row + half) set = some pixel of tile row `row`, columns
 * 8*half..8*half+7, may reach alpha >= 1/255.  Must be a superset of the pixels the blend loop
 * accepts (tests/test_capi.py brute-forces it). */
TS_API uint32_t ts_debug_rowmask(const float* q0_host, const float* q1_host, int tile_x, int tile_y);
/* Test hook (device): rcp_out[i] = MUFU.RCP(x[i]), ex2_out[i] = MUFU.EX2(x[i]) as the blend kernels
 * evaluate them.  Blend-backward relies on rcp(1) == 1 exactly (a pixel a Gaussian does not reach
 * keeps its transmittance through T * rcp(1 - 0)); the GPU tests assert it. */
TS_API int ts_debug_approx(int n, const float* x, float* rcp_out, float* ex2_out, ts_stream_t stream);

/* ---- K6 + K7 in one launch (fused pipeline, single GPU) -----------------------------------------
 * ts_project_bwd (packed-gradient form, TS_PROJ_* flags as there) and ts_sh_bwd (view direction from
 * the mean, colour cotangent = floats 8..10 of the packed row, optional clamp mask) for the same N
 * Gaussians: every CTA builds the SH rows of its Gaussians in shared memory, sends them off as TMA bulk
 * stores and runs the EWA algebra while they drain.  Same results as the two separate calls. */
TS_API int ts_project_sh_bwd(int N, int degree, int K, const float* means3d /*[16B]*/,
                             const float* scales /*[16B]*/, float glob_scale, const float* quats /*[16B]*/,
                             const float* viewmat, const float* projmat, float fx, float fy, float cx,
                             float cy, int img_height, int img_width, int flags, const int32_t* radii,
                             const float* packed_grads /*[16B]*/, const float* opacity_logits /*or NULL*/,
                             const uint8_t* clamp_mask /*or NULL*/, float* v_means3d, float* v_scales,
                             float* v_quats, float* v_opacity_logits /*or NULL*/, float* v_xys_out /*or NULL*/,
                             float* v_dc, float* v_rest, ts_stream_t stream);

/* ---- fused L1 image loss, forward + gradient in one pass (csrc/loss.cu) --------
 * loss[0] = loss_scale * sum_i |img[i] - target[i]|,  grad[i] = grad_scale * sign(img[i] - target[i])
 * (grad may be NULL).  target: float32, or uint8 meaning value / 255 (target_is_u8).  work: a device
 * buffer of ts_l1_loss_work_floats() floats, zeroed ONCE by the caller and then reusable by calls on
 * the same stream.  Deterministic.  [REF scripts/train.py:58-59] */
TS_API int ts_l1_loss_work_floats(void);
TS_API int ts_l1_loss(int64_t n, const float* img /*[16B]*/, const void* target, int target_is_u8,
                      float grad_scale, float loss_scale, float* grad /*[16B] or NULL*/, float* work,
                      float* loss, ts_stream_t stream);

/* ---- data-parallel gradient exchange in packed form -------------------------
 * Instead of all-reducing the finished parameter gradients (236 B per Gaussian at SH degree 3),
 * ranks exchange blend-backward's packed rows (48 B per view and Gaussian): every rank owns a
 * shard of the Gaussians, receives that shard's rows from all views (all-to-all), runs
 * projection-/SH-backward for all views on the shard, and the shard results are all-gathered.
 * ts_dp_prepare: on the rendering rank, makes the packed rows self-contained (exact zeros for
 *   Gaussians culled in this view, SH clamp mask applied to the colour cotangents) and emits
 *   this view's d loss / d xy [N,2] (may be NULL).  recs = the packed raster records.
 * ts_project_bwd_views / ts_sh_bwd_views: backward of a shard of N Gaussians summed over
 *   n_views views and multiplied by out_scale.  cams[n_views][32] (device): floats 0..11 the
 *   3x4 view matrix, 12..27 the 4x4 full projection, 28/29 fx/fy.  packed_grads holds view v's
 *   rows at packed_grads + v * view_stride_floats (+ 12 floats per Gaussian).  flags as
 *   ts_project_bwd (TS_PROJ_DEPTH_CH3: depth cotangent in packed float 11); the SH view
 *   direction is mean - camera translation as with TS_SH_DIRS_FROM_MEANS. */
TS_API int ts_dp_prepare(int N, const int32_t* radii, const uint8_t* clamp_mask /*or NULL*/,
                         const float* recs /*[16B]*/, float* grads /*[16B] in/out*/,
                         float* v_xys /*or NULL*/, ts_stream_t stream);
TS_API int ts_project_bwd_views(int n_views, int N, const float* means3d /*[16B]*/,
                                const float* scales /*[16B]*/, float glob_scale,
                                const float* quats /*[16B]*/, const float* cams,
                                int img_height, int img_width, int flags,
                                const float* packed_grads /*[16B]*/, int64_t view_stride_floats,
                                const float* opacity_logits /*or NULL*/, float out_scale,
                                float* v_means3d /*[16B]*/, float* v_scales /*[16B]*/,
                                float* v_quats /*[16B]*/, float* v_opacity_logits /*or NULL*/,
                                ts_stream_t stream);
TS_API int ts_sh_bwd_views(int n_views, int N, int degree, int K, const float* means,
                           const float* cams, const float* packed_grads /*[16B]*/,
                           int64_t view_stride_floats, float out_scale, float* v_dc,
                           float* v_rest, ts_stream_t stream);

/* ---- the same exchange done by the kernels over NVLink peer memory -----------
 * No collective call inside the step (csrc/peer.cu).  Every rank owns one
 * allocation (ts_peer_alloc: cudaMalloc, zero-filled) and maps the other ranks' allocations through
 * CUDA IPC (ts_peer_ipc_get -> 64-byte handle, exchanged by the host; ts_peer_ipc_open).  Pointer
 * tables (`*_ptrs_host`) are HOST arrays of `world` DEVICE pointers, entry r = the address of that
 * buffer in rank r's allocation as mapped into THIS process.
 * ts_dp_push: after blend-backward, cleans this view's packed rows like ts_dp_prepare and stores
 *   - the 32-byte geometry row {S_x,S_y,S_xx,S_xy | S_yy,v_opacity,v_depth,0} of Gaussian i to its
 *     owner rank i / shard_rows:  geo[owner][rank][i % shard_rows]   (geo = [world][shard_rows][8])
 *   - the colour cotangent (3 floats) to EVERY rank:  rgb[r][rank][i] (rgb = [world][padded_rows][3])
 *   - this view's camera row (32 floats, layout as `cams` above) to every rank: cams[r][rank]
 *   and writes this view's d loss / d xy (v_xys, may be NULL).  shard_rows and padded_rows
 *   (= world * shard_rows) are multiples of 4.
 * ts_peer_barrier: all-to-all barrier through release/acquire flags in peer memory (slot <
 *   ts_peer_barrier_slots(); `epoch` must increase per use of a slot; flags = [slots][8][32] uint32,
 *   ts_peer_flag_bytes()).  mode 1 = signal only (everything this stream did before is visible to a
 *   rank that later waits), 2 = wait only (until every rank has signalled `epoch`), 3 = both.  Signal
 *   and wait may be queued on different streams: a stream that only pushes never blocks on a peer.
 *   A rank that waits longer than timeout_s writes 1 + the missing rank into *err_flag and goes on.
 * ts_sh_bwd_views_rgb: ts_sh_bwd_views reading 12-byte colour rows (rgb[v][i]) instead of packed rows.
 * ts_project_bwd_views_peer: ts_project_bwd_views reading the 32-byte geometry rows and storing the
 *   finished shard gradients to n_dst destinations (tables of n_dst pointers to the shard's first
 *   row in every rank's gradient arrays), starting with destination first_dst. */
TS_API int ts_peer_max_ranks(void);
TS_API int ts_peer_ipc_handle_bytes(void);
TS_API int ts_peer_flag_bytes(void);
TS_API int ts_peer_alloc(int64_t bytes, void** dev_ptr_host);
TS_API int ts_peer_free(void* dev_ptr);
TS_API int ts_peer_ipc_get(void* dev_ptr, void* handle_host);
TS_API int ts_peer_ipc_open(const void* handle_host, void** dev_ptr_host);
TS_API int ts_peer_ipc_close(void* dev_ptr);
TS_API int ts_dp_push(int N, int shard_rows, int padded_rows, int world, int rank, const int32_t* radii,
                      const uint8_t* clamp_mask /*or NULL*/, const float* recs /*[16B]*/,
                      const float* grads /*[16B]*/, const float* cam_row, void* const* geo_ptrs_host,
                      void* const* rgb_ptrs_host, void* const* cam_ptrs_host, float* v_xys /*or NULL*/,
                      int what /*1 geometry rows + camera + v_xys, 2 colour rows, 3 both*/, ts_stream_t stream);
TS_API int ts_peer_barrier(int world, int rank, void* const* flag_ptrs_host, int slot, uint32_t epoch,
                           uint32_t* err_flag, double timeout_s, int mode, ts_stream_t stream);
TS_API int ts_peer_barrier_slots(void);
/* The whole data-parallel backward tail after ts_blend_bwd as one call: a three-stream pipeline over
 * n_pieces (< ts_peer_barrier_slots()) pieces of the rows.  piece_plan_host[c] = {first row, rows, shard
 * rows of the piece, first geometry row of the piece per source rank} (each piece is sharded over the
 * ranks on its own; sizes multiples of 256).  Per piece:
 *   main stream  : ts_dp_push(piece) + signal(slot c)              — never waits for a peer
 *   side stream  : wait(slot c) + ts_sh_bwd_views_rgb(piece)        — local, HBM-bound
 *   side2 stream : ts_project_bwd_views_peer(my shard of the piece) — stores into every rank
 * so the NVLink transfer of piece c+1 runs under the shard backward of piece c; then a barrier in the
 * last slot and the main stream joins.  By default (ts_dp_exchange_split) the geometry rows of a piece
 * are pushed and signalled before its colour rows (two slots per piece): the shard projection-backward,
 * which needs only the geometry, runs under the three times larger colour transfer.  peer_bases_host[world]: every rank's allocation;
 * seg_offsets_host[11]: byte offsets of the segments {flags, err, cams, geo, rgb, g_rest, g_dc, g_means,
 * g_scales, g_quats, g_logit} inside an allocation (tinysplat_b200/parallel.py PeerLayout).  The host
 * side is pure launch logic; issuing it from native code instead of Python keeps it off the critical path. */
/* Debug timeline of ts_dp_exchange_peer: enable, run, synchronize the device, read "label ms" lines
 * (ms since the start of the last exchange; events recorded between the launches on each stream). */
TS_API int ts_dp_exchange_split(int mode /*-1 default/env TINYSPLAT_B200_PEER_SPLIT, 0 off, 1 on*/);
TS_API int ts_dp_exchange_timeline(int enable);
TS_API int ts_dp_exchange_timeline_read(char* buf, int buf_bytes);
TS_API int ts_dp_exchange_peer(int N, int K, int degree, int world, int rank, int n_pieces,
                               const int32_t* piece_plan_host, int padded_rows, const int32_t* radii,
                               const uint8_t* clamp_mask /*or NULL*/, const float* recs /*[16B]*/,
                               const float* grads /*[16B]*/, const float* cam_row, const float* means3d,
                               const float* scales, const float* quats, const float* opacity_logits,
                               void* const* peer_bases_host, const int64_t* seg_offsets_host,
                               int img_height, int img_width, int proj_flags, float out_scale,
                               uint32_t epoch, double timeout_s, float* v_xys /*or NULL*/,
                               ts_stream_t main_stream, ts_stream_t side_stream, ts_stream_t side2_stream);
TS_API int ts_sh_bwd_views_rgb(int n_views, int N, int degree, int K, const float* means,
                               const float* cams, const float* rgb_rows /*[16B]*/,
                               int64_t view_stride_floats, float out_scale, float* v_dc, float* v_rest,
                               ts_stream_t stream);
TS_API int ts_project_bwd_views_peer(int n_views, int N, const float* means3d /*[16B]*/,
                                     const float* scales /*[16B]*/, float glob_scale,
                                     const float* quats /*[16B]*/, const float* cams, int img_height,
                                     int img_width, int flags, const float* geo_rows /*[16B]*/,
                                     int64_t view_stride_floats, const float* opacity_logits /*or NULL*/,
                                     float out_scale, int n_dst, int first_dst,
                                     void* const* v_means_ptrs_host, void* const* v_scales_ptrs_host,
                                     void* const* v_quats_ptrs_host, void* const* v_logit_ptrs_host /*or NULL*/,
                                     ts_stream_t stream);

/* ---- fused multi-tensor Adam step ---------------------------------------
 * One launch updates up to ts_adam_max_tensors() parameter tensors in place, with the exact
 * arithmetic of torch.optim.Adam (no weight decay, no amsgrad), the optimizer the reference
 * trains with [REF scripts/train.py:26; model_gaussian.py:112-120].  The pointer tables,
 * numels[], lrs[] and steps[] (1-based step count AFTER this update) are HOST arrays. */
TS_API int ts_adam_max_tensors(void);
TS_API int ts_adam_step(int num_tensors, float* const* params_host, const float* const* grads_host,
                        float* const* exp_avgs_host, float* const* exp_avg_sqs_host,
                        const int64_t* numels_host, const float* lrs_host, const int64_t* steps_host,
                        double beta1, double beta2, double eps, ts_stream_t stream);

/* ---- fused SSIM (11-tap Gaussian window, 'valid' filtering) ---------------
 * The training loss's (1 - SSIM) term [REF scripts/train.py:60-62; model_gaussian.py:57].
 * X, Y: [B,C,H,W] addressed through ELEMENT strides (host arrays of 4 int64: b, c, h, w), so an
 * [H,W,3] image viewed as [1,3,H,W] is read in place.  win11_host: the 11 window weights (host).
 * ts_ssim_fwd: ssim_sum[B*C] (zeroed by the callee) receives the SUM of the SSIM map per
 * (batch, channel) over the (H-10)x(W-10) valid outputs; dmu/de11/de12 [B,C,H-10,W-10] (all
 * three or all NULL) receive dS/d(mu_x), dS/d(E[x^2]), dS/d(E[xy]) for backward.
 * ts_ssim_bwd: v_X[B,C,H,W] (contiguous) = sum_c v_per_channel[b,c] * d mean_S[b,c] / d X. */
TS_API int ts_ssim_fwd(int B, int C, int H, int W, const float* X, const int64_t* x_strides_host,
                       const float* Y, const int64_t* y_strides_host, const float* win11_host,
                       float C1, float C2, float* ssim_sum, float* dmu, float* de11, float* de12,
                       ts_stream_t stream);
TS_API int ts_ssim_bwd(int B, int C, int H, int W, const float* X, const int64_t* x_strides_host,
                       const float* Y, const int64_t* y_strides_host, const float* win11_host,
                       const float* dmu, const float* de11, const float* de12,
                       const float* v_per_channel, float* v_X, ts_stream_t stream);

/* ---- exact K nearest neighbours (density regularizer) ---------------------
 * Stand-in for pytorch3d.ops.knn_points as called at [REF model_gaussian.py:260,425,519].
 * queries[P1,3], refs[P2,3] -> dists[P1,K] (squared L2, ascending), idx[P1,K] (int64).
 * K in {1,2,4,8,16,32}. */
TS_API int ts_knn_points(int P1, int P2, int K, const float* queries, const float* refs,
                         float* dists, int64_t* idx, ts_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* TINYSPLAT_B200_H */
