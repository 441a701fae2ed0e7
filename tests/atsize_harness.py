"""Helpers of the at-size parity tests (tests/test_gpu_atsize.py): the BASELINE-sized scenes are far
too big for the CPU oracle as a whole, but a WINDOW of tiles only sees the Gaussians whose 3-sigma
tile rectangle reaches it.  These helpers pick those Gaussians with the fp32 oracle projection
(independent of the GPU), render the window with the fp64 oracle's restatement of the reference
adapter [REF tinysplat/splatting/rasterize.py:26-62] and return images and gradients scattered back
to full-size index space.  Everything here is CPU code and is itself tested at small scale in
tests/test_atsize_harness.py."""
import math

import numpy as np
import torch

import oracle

PARAMS = ["means", "scales", "quats", "opacities", "colors_dc", "colors_rest"]

# name -> (N gaussians, W, H, sh degree, depth loss weight, forward only): bench.py's WORKLOADS
CONFIGS = {
    "synthetic_1M_1080p": (1_000_000, 1920, 1080, 3, 0.0, False),
    "synthetic_500k_1080p": (500_000, 1920, 1080, 3, 0.0, False),
    "synthetic_2M_1080p_depthreg": (2_000_000, 1920, 1080, 3, 0.2, False),
    "synthetic_4M_4k_fwd": (4_000_000, 3840, 2160, 3, 0.0, True),
}


def tile_grid(W, H):
    return (W + 15) // 16, (H + 15) // 16


def project_fp32_chunked(sc, cam, W, H, chunk=250_000):
    """fp32 oracle projection of the whole scene, in chunks: (xys, depths, radii) on the CPU."""
    tb = tile_grid(W, H) + (1,)
    V, P = cam.view_matrix.float(), cam.proj_matrix.float()
    outs = []
    with torch.no_grad():
        for s in range(0, sc["means"].shape[0], chunk):
            q = sc["quats"][s:s + chunk].float()
            o = oracle.project_gaussians(sc["means"][s:s + chunk].float(), torch.exp(sc["scales"][s:s + chunk].float()),
                                         1.0, q / q.norm(dim=-1, keepdim=True), V[:3], P @ V, cam.f_x, cam.f_y,
                                         W / 2, H / 2, H, W, tb)
            outs.append(o[:3])
    return tuple(torch.cat([o[i] for o in outs]) for i in range(3))


def window_subset(xys, radii, win, margin_px=3.0):
    """Indices of the Gaussians whose 3-sigma square (+ margin: a ceil() within an ulp of an integer
    may differ between fp32 implementations) can reach tile window win = (tx0, ty0, tx1, ty1)."""
    r = radii.float() + margin_px
    x0, y0, x1, y1 = (16.0 * w for w in win)
    # tile rectangle of the square = tiles floor((c - r)/16) .. floor((c + r)/16): it reaches the window iff
    # the square reaches the window's pixel span widened to whole tiles
    keep = (radii > 0) & (xys[:, 0] + r >= x0 - 16) & (xys[:, 0] - r < x1 + 16) & \
           (xys[:, 1] + r >= y0 - 16) & (xys[:, 1] - r < y1 + 16)
    return torch.nonzero(keep).flatten()


def choose_windows(W, H, tile_counts=None, size=2):
    """Disjoint size x size tile windows: image centre, top-left corner, bottom-right corner (ragged
    last tile row at 1080p) and — given per-tile list lengths — the tile with the longest list."""
    tbx, tby = tile_grid(W, H)
    wins = [(tbx // 2 - 1, tby // 2 - 1, tbx // 2 - 1 + size, tby // 2 - 1 + size), (0, 0, size, size),
            (tbx - size, tby - size, tbx, tby)]
    if tile_counts is not None:
        t = int(torch.as_tensor(tile_counts).argmax())
        tx, ty = t % tbx, t // tbx
        tx0, ty0 = min(tx, tbx - size), min(ty, tby - size)
        cand = (tx0, ty0, tx0 + size, ty0 + size)
        if all(cand[2] <= w[0] or w[2] <= cand[0] or cand[3] <= w[1] or w[3] <= cand[1] for w in wins):
            wins.append(cand)
    return wins


def window_pixels(win, W, H):
    tx0, ty0, tx1, ty1 = win
    return 16 * tx0, 16 * ty0, min(16 * tx1, W), min(16 * ty1, H)


def loss_weights(W, H, wins, seed=0):
    """Full-size loss weights that are non-zero only inside the windows: the loss
    (img * wi).sum() + depth_w * (depth * wd).sum() is then a window-restricted loss."""
    g = torch.Generator().manual_seed(seed)
    wi = torch.zeros(H, W, 3, dtype=torch.float64)
    wd = torch.zeros(H, W, dtype=torch.float64)
    for win in wins:
        x0, y0, x1, y1 = window_pixels(win, W, H)
        wi[y0:y1, x0:x1] = torch.rand(y1 - y0, x1 - x0, 3, generator=g, dtype=torch.float64)
        wd[y0:y1, x0:x1] = torch.rand(y1 - y0, x1 - x0, generator=g, dtype=torch.float64)
    return wi, wd


def oracle_windows(sc, cam, W, H, deg, wins, wi, wd, depth_w, xys32, radii32, want_grads=True,
                   dtype=torch.float64):
    """Oracle (fp64 by default) of every window on the Gaussians that reach it.  Returns (per-window
    list of (img, depth) numpy arrays, grads dict of full-size fp64 tensors or None, union of the
    subsets).  Run a second time with dtype=torch.float32 it calibrates the tolerance: the stated
    algorithm takes DISCRETE decisions (a contribution with alpha < 1/255 is skipped — the cause of the
    first such case seen: ONE pixel of the bottom-right window of the 1M scene, 5.8e-4 —, radius =
    ceil(3 sigma) decides which tiles a Gaussian reaches, equal fp32 depths are ordered by id) and xys is an fp32 tensor of the API (ulp 1.2e-4 px at x ~ 1900),
    so an fp32 evaluation of the same algorithm departs from fp64 by more than the plain rounding
    noise in some windows."""
    N = sc["means"].shape[0]
    imgs = []
    grads = {k: torch.zeros(sc[k].shape, dtype=torch.float64) for k in PARAMS} if want_grads else None
    vxy = torch.zeros(N, 2, dtype=torch.float64) if want_grads else None
    union = torch.zeros(N, dtype=torch.bool)
    for win in wins:
        idx = window_subset(xys32, radii32, win)
        union[idx] = True
        p = {k: (sc[k][idx].to(dtype).clone().requires_grad_(want_grads) if k in PARAMS else sc[k].to(dtype))
             for k in sc}
        img, ex = oracle.render_reference_adapter(p, cam.view_matrix, cam.proj_matrix, cam.f_x, cam.f_y, (W, H), deg,
                                                  tile_window=win)
        x0, y0, x1, y1 = window_pixels(win, W, H)
        if want_grads:
            loss = (img * wi[y0:y1, x0:x1].to(dtype)).sum()
            if depth_w:
                loss = loss + depth_w * (ex["depth"] * wd[y0:y1, x0:x1].to(dtype)).sum()
            loss.backward()
            for k in PARAMS:
                grads[k][idx] += p[k].grad.double()
            if ex["xys"].grad is not None:
                vxy[idx] += ex["xys"].grad.double()
        imgs.append((img.detach().double().numpy(), ex["depth"].detach().double().numpy()))
    if want_grads:
        grads["xys"] = vxy
    return imgs, grads, union


def psnr(a, b):
    mse = float(np.mean((np.asarray(a, np.float64) - np.asarray(b, np.float64)) ** 2))
    return math.inf if mse == 0 else -10.0 * math.log10(mse)


def is_subsequence(sub, full):
    """True when `sub` appears in `full` in the same order (both 1-D integer arrays)."""
    sub, full = np.asarray(sub), np.asarray(full)
    pos = {int(v): i for i, v in enumerate(full)}
    try:
        where = np.array([pos[int(v)] for v in sub], dtype=np.int64)
    except KeyError:
        return False
    return bool(np.all(np.diff(where) > 0)) if where.size > 1 else True


def max_alpha_in_tile(xy, conic, opacity, tile_xy):
    """fp64 maximum over the 256 pixel centres of one tile of opacity * exp(-sigma) (sigma >= 0 only)."""
    tx, ty = tile_xy
    jj, ii = np.meshgrid(16 * tx + np.arange(16) + 0.5, 16 * ty + np.arange(16) + 0.5)
    dx, dy = float(xy[0]) - jj, float(xy[1]) - ii
    a, b, c = (float(v) for v in conic)
    sig = 0.5 * (a * dx * dx + c * dy * dy) + b * dx * dy
    al = np.where(sig >= 0, float(opacity) * np.exp(-sig), 0.0)
    return float(al.max())
