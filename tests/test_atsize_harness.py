"""CPU check of tests/atsize_harness.py at small scale: rendering only the windows on the Gaussians
that reach them must reproduce the crop and the gradients of the full fp64 oracle render under the
same window-restricted loss — the property the at-size GPU tests rely on."""
import numpy as np
import torch

import atsize_harness as ah
import oracle
from tinysplat_b200 import synthetic


def test_windowed_oracle_equals_full_oracle_under_a_window_loss():
    N, W, H, deg, dw = 3000, 320, 200, 3, 0.2          # ragged last tile row (200 = 12.5 tiles)
    cam = synthetic.make_camera(W, H, yaw_deg=2.0)
    sc = synthetic.make_scene(N, W, H, seed=4)
    sc["background"] = torch.tensor([0.2, 0.4, 0.1])
    sc["scales"][:4] += 3.0                              # a few huge Gaussians reach every window
    xys, depths, radii = ah.project_fp32_chunked(sc, cam, W, H, chunk=700)
    tb = ah.tile_grid(W, H)
    tile, gid = oracle.gsplat_oracle.bin_and_sort(xys, depths, radii, tb + (1,))
    counts = torch.bincount(tile, minlength=tb[0] * tb[1])
    wins = ah.choose_windows(W, H, counts)
    assert len(wins) >= 3
    wi, wd = ah.loss_weights(W, H, wins)
    imgs, grads, union = ah.oracle_windows(sc, cam, W, H, deg, wins, wi, wd, dw, xys, radii)

    p = {k: (v.double().clone().requires_grad_(k in ah.PARAMS)) for k, v in sc.items()}
    img, ex = oracle.render_reference_adapter(p, cam.view_matrix, cam.proj_matrix, cam.f_x, cam.f_y, (W, H), deg)
    ((img * wi).sum() + dw * (ex["depth"] * wd).sum()).backward()
    for win, (wimg, wdep) in zip(wins, imgs):
        x0, y0, x1, y1 = ah.window_pixels(win, W, H)
        assert np.abs(wimg - img.detach().numpy()[y0:y1, x0:x1]).max() < 1e-12
        assert np.abs(wdep - ex["depth"].detach().numpy()[y0:y1, x0:x1]).max() < 1e-11
    for k in ah.PARAMS:
        full = p[k].grad
        assert (full[~union] == 0).all(), k              # nothing outside the subsets sees the loss
        assert (grads[k] - full).abs().max().item() <= 1e-9 * max(1.0, full.abs().max().item()), k
    assert (grads["xys"] - ex["xys"].grad).abs().max().item() < 1e-9
    assert 0 < int(union.sum()) < N


def test_subsequence_and_alpha_helpers():
    assert ah.is_subsequence([3, 9, 4], [3, 7, 9, 1, 4])
    assert not ah.is_subsequence([9, 3], [3, 7, 9])
    assert not ah.is_subsequence([5], [3, 7, 9])
    assert ah.is_subsequence([], [1, 2])
    # a Gaussian centred on a pixel centre of tile (1, 0): its maximum alpha there is its opacity
    assert abs(ah.max_alpha_in_tile((16 + 4.5, 7.5), (0.1, 0.0, 0.1), 0.7, (1, 0)) - 0.7) < 1e-12
    assert ah.psnr(np.zeros(4), np.zeros(4)) == float("inf")
