// Compiles the blend kernels (tinysplat_b200/csrc/blend.cu: forward + first-generation backward;
// blend_group.cu: grouped backward) as host code on the fiber SIMT emulator (ts_emu.h) and exposes
// them with the argument lists of ts_blend_fwd / ts_blend_bwd, host pointers instead of device
// pointers.  TEST INFRASTRUCTURE ONLY.
#define TS_HOST_EMU 1
#include "../../tinysplat_b200/csrc/blend.cu"
#include "../../tinysplat_b200/csrc/blend_group.cu"

namespace {
template <int CH>
int run_fwd(int H, int W, int tx, int ty, const int32_t* off, const int32_t* ids, const float* recs,
            const float* bg, float* out_img, float* out_ch3, float* final_T, int32_t* n_contrib, int clamp) {
    return ts_emu::launch(dim3(tx, ty), ts::kBlendThreads, [=]() {
        ts::blend_fwd_kernel<CH>(H, W, tx, off, ids, (const float4*)recs, bg, out_img, out_ch3, final_T,
                                 n_contrib, clamp);
    });
}
template <int CH, int GCH>
int run_bwd(int grouped, int H, int W, int tx, int ty, const int32_t* off, const int32_t* ids, const float* recs,
            const float* bg, const float* final_T, const int32_t* n_contrib, const float* v_img,
            const float* v_ch3, int split, const float* v_alpha, float* grads) {
    if (grouped)
        return ts_emu::launch(dim3(tx, ty), ts::kGThreads, [=]() {
            ts::blend_bwd_group_kernel<CH, GCH>(H, W, tx, off, ids, (const float4*)recs, bg, final_T, n_contrib,
                                                v_img, v_ch3, split, v_alpha, (float4*)grads);
        });
    return ts_emu::launch(dim3(tx, ty), ts::kBlendThreads, [=]() {
        ts::blend_bwd_kernel<CH, GCH>(H, W, tx, off, ids, (const float4*)recs, bg, final_T, n_contrib,
                                      v_img, v_ch3, split, v_alpha, (float4*)grads);
    });
}
}  // namespace

extern "C" {

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
