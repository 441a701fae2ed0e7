// Compiles tinysplat_b200/csrc/binning.cu (tile count -> scan -> emit -> per-tile sort) as host code
// on the fiber SIMT emulator (ts_emu.h) and exposes the four stages with the argument lists of
// ts_bin_count / ts_bin_scan / ts_bin_emit / ts_bin_sort (host pointers; the launch logic below
// mirrors those entry points).  TEST INFRASTRUCTURE ONLY.
#define TS_HOST_EMU 1
#include "../../tinysplat_b200/csrc/binning.cu"

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
        ts::bin_emit_kernel(N, depths, radii, (const float4*)recs, tx, ty, cull, cursors, keys);
    });
}

int emu_bin_sort(int T, const int32_t* offsets, uint64_t* keys, int32_t* ids_sorted, int max_count,
                 int n_big, uint64_t* big_scratch, int32_t* big_counter) {
    if (max_count <= 0) return 0;
    int rc = ts_emu::launch(dim3((T + 7) / 8), 256, [=]() { ts::bin_sort_warp_kernel(T, offsets, keys, ids_sorted); });
    const int bounds[3] = {ts::kWarpSortMax, 2048, ts::kSmemSortCap};
    for (int c = 0; c < 2 && rc == 0; ++c) {
        if (max_count <= bounds[c]) break;
        const int lo = bounds[c], hi = bounds[c + 1];
        rc = ts_emu::launch(dim3(T), ts::kSortThreads, [=]() { ts::bin_sort_kernel(T, offsets, keys, ids_sorted, lo, hi); });
    }
    if (n_big > 0 && rc == 0) {
        int P = 2;
        while (P < max_count) P <<= 1;
        *big_counter = 0;
        rc = ts_emu::launch(dim3(T), 1024, [=]() {
            ts::bin_sort_big_kernel(T, offsets, keys, ids_sorted, ts::kSmemSortCap, P, big_scratch, big_counter);
        });
    }
    return rc;
}

}  // extern "C"
