"""CPU-only work statistics of the blend kernels' culling hierarchy on a synthetic scene.

For a sample of Gaussians of `synthetic_1M_1080p` it counts, per level of the hierarchy
(tile 16x16 -> sub-block 8x4 -> pixel), how many (Gaussian, cell) pairs survive the current
bounding-box test and how many an exact ellipse test would keep.  Used to decide which culling
refinements pay before spending GPU time.  Imports the oracle's projection
only as a measuring instrument; nothing here ships.

    python tests/analysis/cull_stats.py [--n 1000000] [--sample 40000]

Lives under tests/ because it imports the oracle (test infrastructure; the product never does).
"""
import argparse
import math
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle  # noqa: E402
from tinysplat_b200 import synthetic  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--sample", type=int, default=40_000)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    args = ap.parse_args()
    W, H = args.width, args.height
    cam = synthetic.make_camera(W, H)
    sc = synthetic.make_scene(args.n, W, H, seed=0)
    tb = ((W + 15) // 16, (H + 15) // 16, 1)
    with torch.no_grad():
        xys, depths, radii, conics, ntiles, _ = oracle.project_gaussians(
            sc["means"], sc["scales"].exp(), 1.0, torch.nn.functional.normalize(sc["quats"], dim=-1),
            cam.view_matrix[:3], cam.proj_matrix @ cam.view_matrix, cam.f_x, cam.f_y, W / 2, H / 2,
            H, W, tb)
    op = torch.sigmoid(sc["opacities"][:, 0])
    vis = (radii > 0) & (op * 255 >= 1)
    idx = torch.nonzero(vis)[:, 0]
    g = torch.Generator().manual_seed(1)
    idx = idx[torch.randperm(idx.numel(), generator=g)[: args.sample]]
    scale = vis.sum().item() / idx.numel()
    xy = xys[idx].double().numpy()
    con = conics[idx].double().numpy()
    o = op[idx].double().numpy()
    rad = radii[idx].double().numpy()

    a, b, c = con[:, 0], con[:, 1], con[:, 2]
    det = a * c - b * b
    tau = np.log(255.0 * o)
    hx = np.sqrt(2 * tau * c / det)
    hy = np.sqrt(2 * tau * a / det)

    tot = dict(tile_3sigma=0, tile_bbox=0, tile_exact=0, sub_bbox=0, sub_exact=0, sub4_bbox=0,
               sub4_exact=0, pix=0, row8_exact=0, sub4x8_exact=0, row4_in_4x8=0, sub8x8_exact=0, sub8x2_exact=0)
    R = 64  # half window in pixels (Gaussians larger than this are clipped; rare at 6 px mean)
    for k in range(idx.numel()):
        x, y = xy[k]
        # window of pixel indices around the centre, aligned to tiles
        j0 = (int(math.floor(x)) - R) // 16 * 16
        i0 = (int(math.floor(y)) - R) // 16 * 16
        n = 2 * R // 16 + 2
        jj = j0 + np.arange(n * 16)
        ii = i0 + np.arange(n * 16)
        inimg_x = (jj >= 0) & (jj < W)
        inimg_y = (ii >= 0) & (ii < H)
        dx = x - (jj + 0.5)
        dy = y - (ii + 0.5)
        sig = 0.5 * (a[k] * dx[None, :] ** 2 + c[k] * dy[:, None] ** 2) + b[k] * dx[None, :] * dy[:, None]
        lit = (sig <= tau[k]) & (sig >= 0) & inimg_x[None, :] & inimg_y[:, None]
        # bbox test per pixel-centre
        bx = (np.abs(dx) <= hx[k]) & inimg_x
        by = (np.abs(dy) <= hy[k]) & inimg_y
        # 3-sigma box per tile (tile units)
        r = rad[k]
        tx_lo, tx_hi = math.floor((x / 16) - r / 16), math.floor(x / 16 + r / 16 + 1)
        ty_lo, ty_hi = math.floor((y / 16) - r / 16), math.floor(y / 16 + r / 16 + 1)
        tjs = (j0 // 16) + np.arange(n)
        tis = (i0 // 16) + np.arange(n)
        t3x = (tjs >= max(tx_lo, 0)) & (tjs < min(tx_hi, tb[0]))
        t3y = (tis >= max(ty_lo, 0)) & (tis < min(ty_hi, tb[1]))
        tot["tile_3sigma"] += t3x.sum() * t3y.sum()
        tbx_ = bx.reshape(n, 16).any(1) & t3x
        tby_ = by.reshape(n, 16).any(1) & t3y
        tot["tile_bbox"] += tbx_.sum() * tby_.sum()
        tile_ok = np.outer(t3y, t3x)
        lit_t = lit.reshape(n, 16, n, 16).any(axis=(1, 3)) & tile_ok
        tot["tile_exact"] += lit_t.sum()
        # sub-blocks 8 wide x 4 tall, inside tiles that pass the tile test
        tile_pass = np.outer(tby_, tbx_)
        sbx = bx.reshape(n * 2, 8).any(1)
        sby = by.reshape(n * 4, 4).any(1)
        sub_b = np.outer(sby, sbx) & np.repeat(np.repeat(tile_pass, 4, 0), 2, 1)
        tot["sub_bbox"] += sub_b.sum()
        lit_s = lit.reshape(n * 4, 4, n * 2, 8).any(axis=(1, 3)) & np.repeat(np.repeat(tile_ok, 4, 0), 2, 1)
        tot["sub_exact"] += lit_s.sum()
        # 4x4 sub-blocks
        s4x = bx.reshape(n * 4, 4).any(1)
        s4y = by.reshape(n * 4, 4).any(1)
        sub4_b = np.outer(s4y, s4x) & np.repeat(np.repeat(tile_pass, 4, 0), 4, 1)
        tot["sub4_bbox"] += sub4_b.sum()
        lit_4 = lit.reshape(n * 4, 4, n * 4, 4).any(axis=(1, 3)) & np.repeat(np.repeat(tile_ok, 4, 0), 4, 1)
        tot["sub4_exact"] += lit_4.sum()
        # alternative group shapes for the grouped backward: 4 wide x 8 tall, 8 x 8, 8 x 2
        lit_48 = lit.reshape(n * 2, 8, n * 4, 4).any(axis=(1, 3)) & np.repeat(np.repeat(tile_ok, 2, 0), 4, 1)
        tot["sub4x8_exact"] += lit_48.sum()
        lit_88 = lit.reshape(n * 2, 8, n * 2, 8).any(axis=(1, 3)) & np.repeat(np.repeat(tile_ok, 2, 0), 2, 1)
        tot["sub8x8_exact"] += lit_88.sum()
        lit_82 = lit.reshape(n * 8, 2, n * 2, 8).any(axis=(1, 3)) & np.repeat(np.repeat(tile_ok, 8, 0), 2, 1)
        tot["sub8x2_exact"] += lit_82.sum()
        lit_in = lit & np.repeat(np.repeat(tile_ok, 16, 0), 16, 1)
        tot["row4_in_4x8"] += lit_in.reshape(n * 16, n * 4, 4).any(axis=2).sum()
        tot["pix"] += lit_in.sum()
        tot["row8_exact"] += lit_in.reshape(n * 16, n * 2, 8).any(axis=2).sum()

    print(f"N={args.n} visible={int(vis.sum())} sample={idx.numel()} (scaled to the full scene)")
    for k, v in tot.items():
        print(f"  {k:12s} {v * scale / 1e6:9.3f} M")
    print(f"  lanes/iter bbox 8x4 : {tot['pix'] / (32 * tot['sub_bbox']):.3f}")
    print(f"  lanes/iter exact 8x4: {tot['pix'] / (32 * tot['sub_exact']):.3f}")
    print(f"  lanes/iter bbox 4x4 : {tot['pix'] / (16 * tot['sub4_bbox']):.3f}")
    print(f"  lanes/iter exact 4x4: {tot['pix'] / (16 * tot['sub4_exact']):.3f}")
    print(f"  lanes/iter exact row8: {tot['pix'] / (8 * tot['row8_exact']):.3f}")
    print(f"  lanes/iter exact row4: {tot['pix'] / (4 * tot['row4_in_4x8']):.3f}")
    # instruction model of the grouped backward (per warp-iteration: head 32 + rows * 39 + reduce + reds),
    # groups of a warp in lock-step (imbalance ~1.1 for 4 groups, ~1.2 for 8)
    sc = scale / 1e6
    for name, pairs, groups, rows, reduce, imb in (("8x4 (built)", tot["sub_exact"], 4, 4, 46, 1.1),
                                                   ("8x2", tot["sub8x2_exact"], 4, 2, 46, 1.1),
                                                   ("8x8", tot["sub8x8_exact"], 4, 8, 46, 1.1),
                                                   ("4x8 (4-lane groups)", tot["sub4x8_exact"], 8, 8, 36, 1.2)):
        iters = pairs * sc / groups * imb
        print(f"  model {name:22s}: {pairs * sc:6.2f} M pairs, {iters:5.2f} M warp-iterations, "
              f"{iters * (32 + rows * 39 + reduce + 14):6.0f} M warp-instructions")


if __name__ == "__main__":
    main()
