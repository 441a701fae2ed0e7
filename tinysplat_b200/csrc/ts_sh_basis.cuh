// Real spherical-harmonics basis, bands 0..DEG, of a (not necessarily unit) direction — shared by the SH
// kernels (sh.cu) and the fused projection+SH backward (project.cu).
// [behind gsplat.sh.spherical_harmonics, REF tinysplat/splatting/rasterize.py:38,75-81]
#pragma once
#include "ts_common.cuh"

namespace ts {

#define TS_SH_C0 0.28209479177387814f
#define TS_SH_C1 0.4886025119029199f

template <int DEG>
__device__ __forceinline__ void sh_basis(float dx, float dy, float dz, float* b) {
    b[0] = TS_SH_C0;
    if (DEG < 1) return;
    float n2 = dx * dx + dy * dy + dz * dz;
    float inv = rsqrtf(fmaxf(n2, 1e-30f));
    // one Newton step: rsqrtf is approximate, the oracle divides by the exact norm
    inv = inv * (1.5f - 0.5f * n2 * inv * inv);
    float x = dx * inv, y = dy * inv, z = dz * inv;
    b[1] = -TS_SH_C1 * y;
    b[2] = TS_SH_C1 * z;
    b[3] = -TS_SH_C1 * x;
    if (DEG < 2) return;
    float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    b[4] = 1.0925484305920792f * xy;
    b[5] = -1.0925484305920792f * yz;
    b[6] = 0.31539156525252005f * (2.f * zz - xx - yy);
    b[7] = -1.0925484305920792f * xz;
    b[8] = 0.5462742152960396f * (xx - yy);
    if (DEG < 3) return;
    b[9] = -0.5900435899266435f * y * (3.f * xx - yy);
    b[10] = 2.890611442640554f * xy * z;
    b[11] = -0.4570457994644658f * y * (4.f * zz - xx - yy);
    b[12] = 0.3731763325901154f * z * (2.f * zz - 3.f * xx - 3.f * yy);
    b[13] = -0.4570457994644658f * x * (4.f * zz - xx - yy);
    b[14] = 1.445305721320277f * z * (xx - yy);
    b[15] = -0.5900435899266435f * x * (xx - 3.f * yy);
    if (DEG < 4) return;
    b[16] = 2.5033429417967046f * xy * (xx - yy);
    b[17] = -1.7701307697799304f * yz * (3.f * xx - yy);
    b[18] = 0.9461746957575601f * xy * (7.f * zz - 1.f);
    b[19] = -0.6690465435572892f * yz * (7.f * zz - 3.f);
    b[20] = 0.10578554691520431f * (zz * (35.f * zz - 30.f) + 3.f);
    b[21] = -0.6690465435572892f * xz * (7.f * zz - 3.f);
    b[22] = 0.47308734787878004f * (xx - yy) * (7.f * zz - 1.f);
    b[23] = -1.7701307697799304f * xz * (xx - 3.f * yy);
    b[24] = 0.6258357354491761f * (xx * (xx - 3.f * yy) - yy * (3.f * xx - yy));
}

}  // namespace ts
