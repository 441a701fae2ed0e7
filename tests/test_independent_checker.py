"""Cross-check of the two CPU statements of the path: tests/independent_checker.py (sequential
numpy fp64, no shared code with oracle/) against oracle/gsplat_oracle.py on BASELINE config 1
(the committed golden vectors, produced by the oracle) — forward images directly, gradients through
central finite differences of the checker's forward against the oracle's autograd gradients."""
import os

import numpy as np
import pytest
import torch

import independent_checker as chk
import oracle
from tinysplat_b200 import synthetic

GOLD = os.path.join(os.path.dirname(__file__), "golden", "config1.npz")
PARAMS = ["means", "scales", "quats", "opacities", "colors_dc", "colors_rest"]
W = H = 128


def _inputs():
    d = np.load(GOLD)
    p = {k[3:]: d[k].astype(np.float64) for k in d.files if k.startswith("in_")}
    cam = synthetic.make_camera(W, H)
    return d, p, cam


def _loss(p, cam, wi, wd, dir_means):
    rgb, dep = chk.render(p, cam.view_matrix.numpy(), cam.proj_matrix.numpy(), cam.f_x, cam.f_y, W, H, 3,
                          dir_means=dir_means)
    return (rgb * wi).sum() + 0.1 * (dep * wd).sum()


def test_checker_forward_matches_the_oracle_golden_config1():
    d, p, cam = _inputs()
    rgb, dep = chk.render(p, cam.view_matrix.numpy(), cam.proj_matrix.numpy(), cam.f_x, cam.f_y, W, H, 3)
    assert np.abs(rgb - d["img"]).max() < 2e-6          # golden is stored in fp32
    assert np.abs(dep - d["depth"]).max() < 2e-5


def test_checker_finite_differences_match_the_oracle_gradients_config1():
    d, p, cam = _inputs()
    g = torch.Generator().manual_seed(1234)              # the loss weights of tests/golden/make_golden.py
    wi = torch.rand(H, W, 3, generator=g, dtype=torch.float64).numpy()
    wd = torch.rand(H, W, generator=g, dtype=torch.float64).numpy()
    rng = np.random.default_rng(0)
    worst = 0.0
    for k in PARAMS:
        gold = d["v_" + k].astype(np.float64)
        flat = np.abs(gold).ravel()
        # the three largest entries and three random non-zero ones of every tensor
        picks = list(np.argsort(flat)[-3:]) + list(rng.choice(np.nonzero(flat > 1e-3 * flat.max())[0], 3))
        for idx in picks:
            h = 1e-6 * max(1.0, abs(p[k].ravel()[idx]))
            vals = []
            for s in (+1, -1):
                q = {n: v.copy() for n, v in p.items()}
                q[k].ravel()[idx] += s * h
                vals.append(_loss(q, cam, wi, wd, p["means"]))
            fd = (vals[0] - vals[1]) / (2 * h)
            err = abs(fd - gold.ravel()[idx]) / flat.max()
            worst = max(worst, err)
            assert err < 2e-4, (k, int(idx), fd, gold.ravel()[idx])
    print(f"worst FD-vs-autograd error relative to each tensor's max gradient: {worst:.2e}")


def test_checker_window_equals_crop_and_oracle_window():
    """window= renders exactly the crop of the full image, and equals the oracle's tile_window."""
    d, p, cam = _inputs()
    full, fdep = chk.render(p, cam.view_matrix.numpy(), cam.proj_matrix.numpy(), cam.f_x, cam.f_y, W, H, 3)
    win, wdep = chk.render(p, cam.view_matrix.numpy(), cam.proj_matrix.numpy(), cam.f_x, cam.f_y, W, H, 3,
                           window=(32, 48, 80, 96))
    assert np.array_equal(win, full[48:96, 32:80]) and np.array_equal(wdep, fdep[48:96, 32:80])
    pt = {k: torch.from_numpy(v) for k, v in p.items()}
    oimg, oex = oracle.render_reference_adapter(pt, cam.view_matrix, cam.proj_matrix, cam.f_x, cam.f_y,
                                                (W, H), 3, tile_window=(2, 3, 5, 6))
    assert np.abs(oimg.numpy() - win).max() < 1e-12
    assert np.abs(oex["depth"].numpy() - wdep).max() < 1e-11
