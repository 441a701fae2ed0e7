"""Pure-PyTorch, differentiable, fp32/fp64 restatement of the five ``gsplat`` symbols that
tinysplat calls.  TEST INFRASTRUCTURE ONLY — see ``oracle/__init__.py``.

PARITY UNPINNED.  The algorithm lives in a third-party dependency that is absent from
/root/reference: ``gsplat`` (legacy 0.1.x functional API; the reference pins no version —
no requirements/lockfile/submodule).  What anchors this file:

* the reference's own call sites, which fix signatures, shapes and tuple arities:
    - project_gaussians      [REF tinysplat/splatting/rasterize.py:32, :64-73]
    - spherical_harmonics    [REF tinysplat/splatting/rasterize.py:38, :75-81]
    - rasterize_gaussians    [REF tinysplat/splatting/rasterize.py:44, :50, :83-86]
    - num_sh_bases           [REF tinysplat/splatting/rasterize.py:76; model_gaussian.py:71]
    - deg_from_sh            [REF tinysplat/splatting/model_gaussian.py:106]
* the camera conventions the inputs carry [REF tinysplat/scene.py:96-121], the
  NDC->pixel map the reference itself uses [REF tinysplat/scene.py:152-156] and the
  (w,x,y,z) quaternion layout [REF tinysplat/utils.py:41-73];
* the 3DGS paper the reference cites [REF README.md:3] for EWA splatting and front-to-back
  alpha compositing;
* the gsplat-legacy constants below (restated from the published algorithm; hypotheses per
  SURVEY.md section 8c).  Because nothing in the reference pins them, "parity" everywhere in
  this repo means "parity with THIS stated oracle".

Forward semantics are the sequential per-pixel algorithm; backward is whatever autograd
derives from it (i.e. the exact a.e. derivative of the stated forward).  That is also the
contract of the CUDA kernels: they implement the exact derivative of this forward,
including a zero gradient through the alpha clamp and through the view-frustum clamp of
the EWA Jacobian.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch
from torch import Tensor

# ---- constants of the stated algorithm (single source of truth for tests) ----------------
BLOCK = 16            # tile edge in pixels [REF rasterize.py:19-20 BLOCK_X/BLOCK_Y]
COV2D_BLUR = 0.3      # added to the diagonal of the projected 2D covariance
RADIUS_SIGMA = 3.0    # radius = ceil(3 * sqrt(lambda_max))
EIG_FLOOR = 0.1       # floor under the eigenvalue discriminant
FOV_CLAMP = 1.3       # view-frustum clamp of t.x/t.z, t.y/t.z in the EWA Jacobian
ALPHA_MAX = 0.999     # alpha = min(ALPHA_MAX, opacity * exp(-sigma))
ALPHA_MIN = 1.0 / 255.0   # contributions below this are skipped
T_STOP = 1e-4         # a pixel stops before the Gaussian that would bring T to <= T_STOP
NEAR_CLIP = 0.01      # default clip_thresh: Gaussians with z_cam <= this are culled
W_EPS = 1e-6          # added to the homogeneous w before the perspective divide
PIX_CENTER = 0.5      # pixel (i, j) is sampled at (j + 0.5, i + 0.5)

SH_C0 = 0.28209479177387814
SH_C1 = 0.4886025119029199
SH_C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005,
         -1.0925484305920792, 0.5462742152960396)
SH_C3 = (-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154,
         -0.4570457994644658, 1.445305721320277, -0.5900435899266435)
SH_C4 = (2.5033429417967046, -1.7701307697799304, 0.9461746957575601, -0.6690465435572892,
         0.10578554691520431, -0.6690465435572892, 0.47308734787878004, -1.7701307697799304,
         0.6258357354491761)


# ---- SH helpers ---------------------------------------------------------------------------
def num_sh_bases(degree: int) -> int:
    """(degree+1)^2 for degree 0..4  [REF model_gaussian.py:71-72 sizes colors with it]."""
    if degree < 0 or degree > 4:
        raise ValueError(f"SH degree must be in 0..4, got {degree}")
    return (degree + 1) ** 2


def deg_from_sh(num_bases: int) -> int:
    """Inverse of num_sh_bases  [REF model_gaussian.py:106]."""
    table = {1: 0, 4: 1, 9: 2, 16: 3, 25: 4}
    if num_bases not in table:
        raise ValueError(f"Invalid number of SH bases: {num_bases}")
    return table[num_bases]


def sh_basis(degree: int, dirs: Tensor, num_bases: int) -> Tensor:
    """Real SH basis values [N, num_bases]; columns above (degree+1)^2 are zero.

    Directions are normalised inside (the caller already passes unit vectors
    [REF rasterize.py:78-79]) and receive no gradient (dirs are detached)."""
    d = dirs.detach()
    n = d.norm(dim=-1, keepdim=True).clamp_min(1e-30)
    x, y, z = (d / n).unbind(-1)
    cols = [torch.full_like(x, SH_C0)]
    if degree >= 1:
        cols += [-SH_C1 * y, SH_C1 * z, -SH_C1 * x]
    if degree >= 2:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        cols += [SH_C2[0] * xy, SH_C2[1] * yz, SH_C2[2] * (2 * zz - xx - yy),
                 SH_C2[3] * xz, SH_C2[4] * (xx - yy)]
    if degree >= 3:
        cols += [SH_C3[0] * y * (3 * xx - yy), SH_C3[1] * xy * z,
                 SH_C3[2] * y * (4 * zz - xx - yy),
                 SH_C3[3] * z * (2 * zz - 3 * xx - 3 * yy),
                 SH_C3[4] * x * (4 * zz - xx - yy), SH_C3[5] * z * (xx - yy),
                 SH_C3[6] * x * (xx - 3 * yy)]
    if degree >= 4:
        cols += [SH_C4[0] * xy * (xx - yy), SH_C4[1] * yz * (3 * xx - yy),
                 SH_C4[2] * xy * (7 * zz - 1), SH_C4[3] * yz * (7 * zz - 3),
                 SH_C4[4] * (zz * (35 * zz - 30) + 3), SH_C4[5] * xz * (7 * zz - 3),
                 SH_C4[6] * (xx - yy) * (7 * zz - 1), SH_C4[7] * xz * (xx - 3 * yy),
                 SH_C4[8] * (xx * (xx - 3 * yy) - yy * (3 * xx - yy))]
    B = torch.stack(cols, dim=-1)
    if B.shape[-1] < num_bases:
        B = torch.cat([B, B.new_zeros(B.shape[0], num_bases - B.shape[-1])], dim=-1)
    return B


def spherical_harmonics(degrees_to_use: int, viewdirs: Tensor, coeffs: Tensor) -> Tensor:
    """[N,3] colour = sum_k basis_k(dir) * coeffs[:, k, :]  [REF rasterize.py:38,81].

    coeffs is [N, K, 3] with K >= (degrees_to_use+1)^2; only the first (deg+1)^2 bases are
    used (the reference always passes K=16, even when the active degree is lower
    [REF rasterize.py:75-81]).  The +0.5 offset and clamp are the caller's
    [REF rasterize.py:39]."""
    K = coeffs.shape[-2]
    if K < num_sh_bases(degrees_to_use):
        raise ValueError("coeffs has fewer bases than degrees_to_use needs")
    B = sh_basis(degrees_to_use, viewdirs.to(coeffs.dtype), K)
    return (B[:, :, None] * coeffs).sum(dim=1)


# ---- projection ---------------------------------------------------------------------------
def quat_to_rotmat(q: Tensor) -> Tensor:
    """(w,x,y,z) -> R, used as given (the caller normalises in torch
    [REF rasterize.py:73]; layout per [REF utils.py:44])."""
    w, x, y, z = q.unbind(-1)
    return torch.stack([
        torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)], -1),
        torch.stack([2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)], -1),
        torch.stack([2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], -1),
    ], dim=-2)


def tile_bbox(xys: Tensor, radii: Tensor, tile_bounds: Sequence[int]):
    """Tile rectangle [min, max) hit by the square of half-edge `radius` around `xy`."""
    tbx, tby = int(tile_bounds[0]), int(tile_bounds[1])
    # arithmetic is done in the dtype of xys on purpose: the fp32 path then rounds exactly
    # like the fp32 kernels (x/16 and radius/16 are exact, the add/sub round once)
    dt = xys.dtype
    c = xys.detach() / BLOCK
    r = radii.to(dt)[:, None] / BLOCK
    lo = torch.floor((c - r).clamp(-1e9, 1e9))
    hi = torch.floor((c + r + 1).clamp(-1e9, 1e9))
    bound = torch.tensor([tbx, tby], dtype=dt)
    lo = torch.minimum(lo.clamp_min(0), bound).to(torch.int64)
    hi = torch.minimum(hi.clamp_min(0), bound).to(torch.int64)
    return lo, hi


def project_gaussians(means3d: Tensor, scales: Tensor, glob_scale: float, quats: Tensor,
                      viewmat: Tensor, projmat: Tensor, fx: float, fy: float, cx: float,
                      cy: float, img_height: int, img_width: int,
                      tile_bounds: Sequence[int], clip_thresh: float = NEAR_CLIP):
    """EWA projection  [REF rasterize.py:32 call, :64-73 argument marshalling].

    Returns the 6-tuple the reference unpacks: (xys[N,2], depths[N], radii[N] int32,
    conics[N,3], num_tiles_hit[N] int32, cov3d[N,6]).  Culled Gaussians (behind the near
    plane, singular 2D covariance, or zero tile area) have every output zero (cov3d is kept
    for those past the near plane)."""
    dt = means3d.dtype
    viewmat = viewmat.to(dt)
    projmat = projmat.to(dt)
    fx, fy, cx, cy = float(fx), float(fy), float(cx), float(cy)
    N = means3d.shape[0]
    Rv, tv = viewmat[:3, :3], viewmat[:3, 3]
    t = means3d @ Rv.T + tv
    z = t[:, 2]
    near_ok = z.detach() > clip_thresh
    zs = torch.where(near_ok, z, torch.ones_like(z))

    # 3D covariance from scale and rotation
    R = quat_to_rotmat(quats)
    M = R * (glob_scale * scales)[:, None, :]
    cov3d = M @ M.transpose(1, 2)

    # EWA: J (with frustum clamp) * view rotation
    limx = FOV_CLAMP * 0.5 * img_width / fx
    limy = FOV_CLAMP * 0.5 * img_height / fy
    txc = zs * torch.clamp(t[:, 0] / zs, -limx, limx)
    tyc = zs * torch.clamp(t[:, 1] / zs, -limy, limy)
    rz = 1.0 / zs
    rz2 = rz * rz
    zero = torch.zeros_like(rz)
    J = torch.stack([
        torch.stack([fx * rz, zero, -fx * txc * rz2], -1),
        torch.stack([zero, fy * rz, -fy * tyc * rz2], -1),
    ], dim=-2)
    T = J @ Rv
    cov2d = T @ cov3d @ T.transpose(1, 2)
    a = cov2d[:, 0, 0] + COV2D_BLUR
    b = cov2d[:, 0, 1]
    c = cov2d[:, 1, 1] + COV2D_BLUR
    det = a * c - b * b
    det_ok = det.detach() != 0
    dets = torch.where(det_ok, det, torch.ones_like(det))
    conic = torch.stack([c / dets, -b / dets, a / dets], dim=-1)
    with torch.no_grad():
        mid = 0.5 * (a + c)
        disc = torch.sqrt(torch.clamp(mid * mid - det, min=EIG_FLOOR))
        lam = torch.maximum(mid + disc, mid - disc)
        radius = torch.ceil(RADIUS_SIGMA * torch.sqrt(lam))
        radius = torch.nan_to_num(radius, nan=0.0, posinf=2.0 ** 30).clamp(0, 2.0 ** 30)

    # perspective projection to pixels [REF scene.py:152-156 uses the same -0.5 convention]
    p_hom = means3d @ projmat[:, :3].T + projmat[:, 3]
    w = p_hom[:, 3] + W_EPS
    w_ok = near_ok & (w.detach() != 0)
    ws = torch.where(w_ok, w, torch.ones_like(w))
    xy = torch.stack([0.5 * img_width * p_hom[:, 0] / ws + cx - 0.5,
                      0.5 * img_height * p_hom[:, 1] / ws + cy - 0.5], dim=-1)

    with torch.no_grad():
        xy_safe = torch.nan_to_num(xy.detach(), nan=0.0, posinf=1e9, neginf=-1e9)
        lo, hi = tile_bbox(xy_safe, radius, tile_bounds)
        area = (hi[:, 0] - lo[:, 0]) * (hi[:, 1] - lo[:, 1])
        # a NaN covariance (zero-norm / NaN quaternion, NaN log-scale) passes det != 0 and has
        # radius 0: culled like every other Gaussian that cannot reach a tile
        ok = near_ok & det_ok & w_ok & (area > 0) & (radius > 0) & (det.detach() == det.detach())

    zt = torch.zeros((), dtype=dt)
    xys = torch.where(ok[:, None], xy, zt)
    depths = torch.where(ok, z, zt)
    conics = torch.where(ok[:, None], conic, zt)
    radii = torch.where(ok, radius, torch.zeros_like(radius)).to(torch.int32)
    num_tiles_hit = torch.where(ok, area, torch.zeros_like(area)).to(torch.int32)
    iu = torch.triu_indices(3, 3)
    cov3d_triu = torch.where(near_ok[:, None], cov3d[:, iu[0], iu[1]], zt)
    return xys, depths, radii, conics, num_tiles_hit, cov3d_triu


# ---- binning + blending --------------------------------------------------------------------
def bin_and_sort(xys: Tensor, depths: Tensor, radii: Tensor, tile_bounds: Sequence[int],
                 tile_window: Optional[Tuple[int, int, int, int]] = None):
    """(tile_id, gaussian_id) intersections sorted by tile, then depth, then gaussian id
    (= a stable sort of the emission order on the key tile<<32 | depth).

    tile_window = (tx0, ty0, tx1, ty1) keeps only tiles in [tx0,tx1) x [ty0,ty1)."""
    tbx, tby = int(tile_bounds[0]), int(tile_bounds[1])
    lo, hi = tile_bbox(xys, radii, tile_bounds)
    if tile_window is not None:
        tx0, ty0, tx1, ty1 = tile_window
        wlo = torch.tensor([tx0, ty0])
        whi = torch.tensor([tx1, ty1])
        lo = torch.maximum(lo, wlo)
        hi = torch.minimum(hi, whi)
    span = (hi - lo).clamp_min(0)
    cnt = span[:, 0] * span[:, 1]
    cnt = torch.where(radii > 0, cnt, torch.zeros_like(cnt))
    gid = torch.repeat_interleave(torch.arange(xys.shape[0]), cnt)
    first = torch.cumsum(cnt, 0) - cnt
    k = torch.arange(gid.shape[0]) - first[gid]
    w = span[gid, 0].clamp_min(1)
    tile = (lo[gid, 1] + k // w) * tbx + (lo[gid, 0] + k % w)
    o1 = torch.argsort(depths.detach()[gid], stable=True)
    gid, tile = gid[o1], tile[o1]
    o2 = torch.argsort(tile, stable=True)
    return tile[o2], gid[o2]


def _blend_tile(xy, con, op, col, px, py, bg):
    """Dense [P pixels x n Gaussians] front-to-back compositing of one tile's sorted list."""
    dx = xy[None, :, 0] - px[:, None]
    dy = xy[None, :, 1] - py[:, None]
    sigma = 0.5 * (con[None, :, 0] * dx * dx + con[None, :, 2] * dy * dy) \
        + con[None, :, 1] * dx * dy
    alpha = torch.clamp(op[None, :] * torch.exp(-sigma), max=ALPHA_MAX)
    valid = (sigma.detach() >= 0) & (alpha.detach() >= ALPHA_MIN)
    one_m = torch.where(valid, 1.0 - alpha, torch.ones_like(alpha))
    T_after = torch.cumprod(one_m, dim=1)
    T_before = torch.cat([torch.ones_like(T_after[:, :1]), T_after[:, :-1]], dim=1)
    included = valid & (T_after.detach() > T_STOP)
    wgt = torch.where(included, alpha * T_before, torch.zeros_like(alpha))
    T_final = torch.where(included, 1.0 - alpha, torch.ones_like(alpha)).prod(dim=1)
    out = wgt @ col + T_final[:, None] * bg[None, :]
    n_contrib = torch.where(included.any(dim=1),
                            included.shape[1] - included.flip(1).to(torch.int64).argmax(dim=1),
                            torch.zeros(included.shape[0], dtype=torch.int64))
    return out, T_final, n_contrib


def rasterize_gaussians(xys: Tensor, depths: Tensor, radii: Tensor, conics: Tensor,
                        num_tiles_hit: Tensor, colors: Tensor, opacity: Tensor,
                        img_height: int, img_width: int, background: Optional[Tensor] = None,
                        tile_window: Optional[Tuple[int, int, int, int]] = None,
                        return_aux: bool = False):
    """Tile-binned, depth-sorted alpha compositing  [REF rasterize.py:44, :50, :83-86].

    Returns the 2-tuple the reference unpacks: (out_img[H,W,C], out_alpha[H,W]) with
    out_alpha = 1 - final transmittance.  With tile_window the outputs cover only that window
    (bounded CPU-baseline sample of a large workload)."""
    dt = colors.dtype
    C = colors.shape[-1]
    H, W = int(img_height), int(img_width)
    tb = ((W + BLOCK - 1) // BLOCK, (H + BLOCK - 1) // BLOCK, 1)
    if background is None:
        background = torch.ones(C, dtype=dt)
    bg = background.to(dt)
    op = opacity.reshape(-1)
    tile, gid = bin_and_sort(xys, depths, radii, tb, tile_window)
    tx0, ty0, tx1, ty1 = tile_window if tile_window is not None else (0, 0, tb[0], tb[1])
    edges = torch.searchsorted(tile, torch.arange(tb[0] * tb[1] + 1))
    rows, rows_T, rows_n = [], [], []
    for ty in range(ty0, ty1):
        y0, y1 = ty * BLOCK, min((ty + 1) * BLOCK, H)
        blocks, blocks_T, blocks_n = [], [], []
        for tx in range(tx0, tx1):
            x0, x1 = tx * BLOCK, min((tx + 1) * BLOCK, W)
            th, tw = y1 - y0, x1 - x0
            t_id = ty * tb[0] + tx
            g = gid[edges[t_id]:edges[t_id + 1]]
            jj, ii = torch.meshgrid(torch.arange(x0, x1), torch.arange(y0, y1), indexing="xy")
            px = (jj.reshape(-1).to(dt) + PIX_CENTER)
            py = (ii.reshape(-1).to(dt) + PIX_CENTER)
            if g.numel() == 0:
                out = bg[None, :].expand(th * tw, C)
                Tf = torch.ones(th * tw, dtype=dt)
                nc = torch.zeros(th * tw, dtype=torch.int64)
            else:
                out, Tf, nc = _blend_tile(xys[g], conics[g], op[g], colors[g], px, py, bg)
            blocks.append(out.reshape(th, tw, C))
            blocks_T.append(Tf.reshape(th, tw))
            blocks_n.append(nc.reshape(th, tw))
        rows.append(torch.cat(blocks, dim=1))
        rows_T.append(torch.cat(blocks_T, dim=1))
        rows_n.append(torch.cat(blocks_n, dim=1))
    img = torch.cat(rows, dim=0)
    T_img = torch.cat(rows_T, dim=0)
    if return_aux:
        return img, 1.0 - T_img, {"n_contrib": torch.cat(rows_n, dim=0),
                                  "tile": tile, "gid": gid, "edges": edges}
    return img, 1.0 - T_img


# ---- the reference adapter's arithmetic, restated with oracle ops --------------------------
def render_reference_adapter(params: dict, view_matrix: Tensor, proj_matrix: Tensor,
                             fx: float, fy: float, dims: Tuple[int, int], sh_degree: int,
                             tile_window=None):
    """What GaussianRasterizer.__call__ computes  [REF rasterize.py:26-62], with the oracle
    standing in for gsplat.  params: means[N,3], scales[N,3] (log), quats[N,4],
    opacities[N,1] (logit), colors_dc[N,3], colors_rest[N,K-1,3], background[3].
    dims = (width, height).  Returns (rgb[H,W,3] clamped to <=1, extras)."""
    W, H = dims
    dt = params["means"].dtype
    V = view_matrix.to(dt)
    P = proj_matrix.to(dt)
    tb = ((W + BLOCK - 1) // BLOCK, (H + BLOCK - 1) // BLOCK, 1)
    q = params["quats"]
    xys, depths, radii, conics, ntiles, _ = project_gaussians(
        params["means"], torch.exp(params["scales"]), 1.0, q / q.norm(dim=-1, keepdim=True),
        V[:3, :], P @ V, fx, fy, W / 2, H / 2, H, W, tb)
    if xys.requires_grad:
        xys.retain_grad()
    dirs = params["means"] - V[:3, 3]
    dirs = dirs / dirs.norm(dim=-1, keepdim=True)
    coeffs = torch.cat([params["colors_dc"][:, None, :], params["colors_rest"]], dim=1)
    rgbs = torch.clamp(spherical_harmonics(sh_degree, dirs, coeffs) + 0.5, min=0.0)
    opac = torch.sigmoid(params["opacities"])
    bg = params["background"].to(dt)
    img, _ = rasterize_gaussians(xys, depths, radii, conics, ntiles, rgbs, opac, H, W, bg,
                                 tile_window=tile_window)
    img = torch.clamp(img, max=1.0)
    dimg, _ = rasterize_gaussians(xys, depths, radii, conics, ntiles,
                                  depths[:, None].repeat(1, 3), opac, H, W, bg,
                                  tile_window=tile_window)
    extras = {"depth": dimg[:, :, 0], "radii": radii, "xys": xys,
              "camera": {"height": H, "width": W}}
    return img, extras
