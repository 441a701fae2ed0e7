"""Pins the oracle (CPU): fp64 gradcheck, closed forms, invariances, golden vectors.
The reference ships no fixtures for this path (PARITY UNPINNED) — these are what anchor it."""
import math
import os

import numpy as np
import pytest
import torch

import oracle
from oracle import gsplat_oracle as go
from tinysplat_b200 import synthetic

GOLD = os.path.join(os.path.dirname(__file__), "golden", "config1.npz")


def _cam(W, H):
    return synthetic.make_camera(W, H)


def _project(p, cam, W, H, dtype=torch.float64):
    tb = ((W + 15) // 16, (H + 15) // 16, 1)
    V, P = cam.view_matrix.to(dtype), cam.proj_matrix.to(dtype)
    q = p["quats"]
    return oracle.project_gaussians(p["means"], torch.exp(p["scales"]), 1.0,
                                    q / q.norm(dim=-1, keepdim=True), V[:3], P @ V, cam.f_x, cam.f_y,
                                    W / 2, H / 2, H, W, tb)


def test_sh_tables():
    assert [oracle.num_sh_bases(d) for d in range(5)] == [1, 4, 9, 16, 25]
    assert [oracle.deg_from_sh(n) for n in (1, 4, 9, 16, 25)] == [0, 1, 2, 3, 4]
    with pytest.raises(ValueError):
        oracle.deg_from_sh(5)


def test_sh_degree0_is_constant_and_orthonormal_bands():
    g = torch.Generator().manual_seed(0)
    d = torch.randn(4000, 3, generator=g, dtype=torch.float64)
    B = go.sh_basis(4, d, 25)
    # Monte-Carlo orthonormality on the sphere: (4 pi / n) B^T B ~ I
    G = (4 * math.pi / d.shape[0]) * B.T @ B
    assert (G - torch.eye(25, dtype=torch.float64)).abs().max() < 0.15
    c = torch.randn(4000, 25, 3, generator=g, dtype=torch.float64)
    out0 = oracle.spherical_harmonics(0, d, c)
    assert torch.allclose(out0, go.SH_C0 * c[:, 0, :])


def test_sh_gradcheck():
    g = torch.Generator().manual_seed(1)
    d = torch.randn(5, 3, generator=g, dtype=torch.float64)
    c = torch.randn(5, 16, 3, generator=g, dtype=torch.float64, requires_grad=True)
    assert torch.autograd.gradcheck(lambda cc: oracle.spherical_harmonics(3, d, cc), (c,))
    # bases above the active degree get exactly zero gradient
    oracle.spherical_harmonics(1, d, c).sum().backward()
    assert c.grad[:, 4:, :].abs().max() == 0


def test_project_gradcheck_fp64():
    W = H = 32
    cam = _cam(W, H)
    sc = synthetic.make_scene(6, W, H, seed=3, dtype=torch.float64)
    names = ["means", "scales", "quats"]
    inputs = tuple(sc[n].clone().requires_grad_(True) for n in names)

    def f(means, scales, quats):
        xys, depths, radii, conics, nt, _ = _project({"means": means, "scales": scales, "quats": quats},
                                                     cam, W, H)
        assert (radii > 0).all()
        return xys, depths, conics

    assert torch.autograd.gradcheck(f, inputs, eps=1e-7, atol=1e-5, rtol=1e-4)


def test_rasterize_gradcheck_fp64():
    W = H = 16
    g = torch.Generator().manual_seed(5)
    N = 5
    xys = (torch.rand(N, 2, generator=g, dtype=torch.float64) * 12 + 2).requires_grad_(True)
    depths = torch.rand(N, generator=g, dtype=torch.float64) + 1
    radii = torch.full((N,), 10, dtype=torch.int32)
    A = torch.randn(N, 2, 2, generator=g, dtype=torch.float64)
    cov = A @ A.transpose(1, 2) + 6 * torch.eye(2, dtype=torch.float64)
    inv = torch.linalg.inv(cov)
    conics = torch.stack([inv[:, 0, 0], inv[:, 0, 1], inv[:, 1, 1]], -1).requires_grad_(True)
    colors = torch.rand(N, 3, generator=g, dtype=torch.float64).requires_grad_(True)
    opac = (torch.rand(N, 1, generator=g, dtype=torch.float64) * 0.6 + 0.2).requires_grad_(True)
    bg = torch.tensor([0.1, 0.2, 0.3], dtype=torch.float64)
    nt = torch.ones(N, dtype=torch.int32)

    def f(xys, conics, colors, opac):
        return oracle.rasterize_gaussians(xys, depths, radii, conics, nt, colors, opac, H, W, bg)

    assert torch.autograd.gradcheck(f, (xys, conics, colors, opac), eps=1e-7, atol=1e-6, rtol=1e-4)


def _one_gaussian(x, y, s, o, c, W=32, H=32, depth=1.0):
    xys = torch.tensor([[x, y]], dtype=torch.float64)
    con = torch.tensor([[1 / s ** 2, 0.0, 1 / s ** 2]], dtype=torch.float64)
    return (xys, torch.tensor([depth], dtype=torch.float64), torch.tensor([40], dtype=torch.int32), con,
            torch.tensor([4], dtype=torch.int32), torch.tensor([c], dtype=torch.float64),
            torch.tensor([[o]], dtype=torch.float64), H, W)


def test_closed_form_single_isotropic_gaussian():
    s, o, c = 4.0, 0.8, [0.9, 0.5, 0.1]
    bg = torch.tensor([0.3, 0.3, 0.3], dtype=torch.float64)
    img, alpha = oracle.rasterize_gaussians(*_one_gaussian(16.0, 16.0, s, o, c), bg)
    jj, ii = torch.meshgrid(torch.arange(32.), torch.arange(32.), indexing="xy")
    r2 = (jj + 0.5 - 16) ** 2 + (ii + 0.5 - 16) ** 2
    a = torch.clamp(o * torch.exp(-r2.double() / (2 * s * s)), max=0.999)
    a = torch.where(a >= 1 / 255, a, torch.zeros_like(a))
    want = a[..., None] * torch.tensor(c, dtype=torch.float64) + (1 - a)[..., None] * bg
    assert torch.allclose(img, want, atol=1e-12)
    assert torch.allclose(alpha, a, atol=1e-12)


def test_closed_form_two_coincident_gaussians_depth_order():
    bg = torch.zeros(3, dtype=torch.float64)
    a1 = _one_gaussian(8.5, 8.5, 3.0, 0.6, [1.0, 0.0, 0.0], 16, 16, depth=2.0)
    a2 = _one_gaussian(8.5, 8.5, 3.0, 0.5, [0.0, 1.0, 0.0], 16, 16, depth=1.0)   # nearer
    cat = lambda i: torch.cat([a1[i], a2[i]])
    img, _ = oracle.rasterize_gaussians(cat(0), cat(1), cat(2), cat(3), cat(4), cat(5), cat(6), 16, 16, bg)
    # pixel (8,8) is exactly at the centre: alpha_i = o_i
    px = img[8, 8]
    assert torch.allclose(px, torch.tensor([0.6 * (1 - 0.5), 0.5, 0.0], dtype=torch.float64), atol=1e-12)


def test_behind_camera_contributes_nothing_and_gets_zero_grad():
    W = H = 32
    cam = _cam(W, H)
    sc = synthetic.make_scene(8, W, H, seed=2, dtype=torch.float64)
    sc["means"][0, 2] = -3.0          # behind the camera
    sc["means"][1, 2] = 0.005         # inside the near clip
    p = {k: v.clone().requires_grad_(True) if k != "background" else v for k, v in sc.items()}
    img, ex = oracle.render_reference_adapter(p, cam.view_matrix, cam.proj_matrix, cam.f_x, cam.f_y, (W, H), 3)
    assert ex["radii"][0] == 0 and ex["radii"][1] == 0
    (img.sum() + ex["depth"].sum()).backward()
    for k in ("means", "scales", "quats"):
        assert p[k].grad[:2].abs().max() == 0
        assert torch.isfinite(p[k].grad).all()


def test_permutation_and_depth_shift_invariance():
    W = H = 48
    cam = _cam(W, H)
    sc = synthetic.make_scene(40, W, H, seed=7, dtype=torch.float64)
    img, ex = oracle.render_reference_adapter(sc, cam.view_matrix, cam.proj_matrix, cam.f_x, cam.f_y, (W, H), 3)
    perm = torch.randperm(40, generator=torch.Generator().manual_seed(0))
    sc2 = {k: (v[perm] if k != "background" else v) for k, v in sc.items()}
    img2, _ = oracle.render_reference_adapter(sc2, cam.view_matrix, cam.proj_matrix, cam.f_x, cam.f_y, (W, H), 3)
    assert torch.allclose(img, img2, atol=1e-12)
    # shifting every depth key by a constant leaves the RGB image unchanged
    xys, depths, radii, conics, nt, _ = _project(sc, cam, W, H)
    col = torch.rand(40, 3, dtype=torch.float64)
    op = torch.sigmoid(sc["opacities"])
    a, _ = oracle.rasterize_gaussians(xys, depths, radii, conics, nt, col, op, H, W, None)
    b, _ = oracle.rasterize_gaussians(xys, depths + 5.0, radii, conics, nt, col, op, H, W, None)
    assert torch.equal(a, b)


def test_tile_window_matches_full_render():
    W, H = 80, 64
    cam = _cam(W, H)
    sc = synthetic.make_scene(60, W, H, seed=9, dtype=torch.float64)
    full, _ = oracle.render_reference_adapter(sc, cam.view_matrix, cam.proj_matrix, cam.f_x, cam.f_y, (W, H), 2)
    win, _ = oracle.render_reference_adapter(sc, cam.view_matrix, cam.proj_matrix, cam.f_x, cam.f_y, (W, H), 2,
                                             tile_window=(1, 1, 4, 3))
    assert torch.equal(win, full[16:48, 16:64])


def test_ragged_image_and_empty_scene():
    W, H = 37, 21           # not multiples of 16
    cam = _cam(W, H)
    sc = synthetic.make_scene(30, W, H, seed=11, dtype=torch.float64)
    img, ex = oracle.render_reference_adapter(sc, cam.view_matrix, cam.proj_matrix, cam.f_x, cam.f_y, (W, H), 1)
    assert img.shape == (H, W, 3) and ex["depth"].shape == (H, W)
    empty = {k: (v[:0] if k != "background" else v) for k, v in sc.items()}
    empty["background"] = torch.tensor([0.25, 0.5, 0.75], dtype=torch.float64)
    img0, ex0 = oracle.render_reference_adapter(empty, cam.view_matrix, cam.proj_matrix, cam.f_x, cam.f_y, (W, H), 1)
    assert torch.allclose(img0, empty["background"].expand(H, W, 3))


def test_golden_config1_fp32_oracle_matches_committed_fp64_vectors():
    """The committed vectors came from the fp64 oracle; the fp32 oracle must agree to the
    calibrated fp32-vs-fp64 gap (tests/golden/make_golden.py prints it: <= ~6e-6 relative)."""
    gold = np.load(GOLD)
    sc = {k[3:]: torch.from_numpy(gold[k]) for k in gold.files if k.startswith("in_")}
    cam = _cam(128, 128)
    p = {k: v.clone().requires_grad_(k != "background") for k, v in sc.items()}
    img, ex = oracle.render_reference_adapter(p, cam.view_matrix, cam.proj_matrix, cam.f_x, cam.f_y, (128, 128), 3)
    assert torch.equal(ex["radii"], torch.from_numpy(gold["radii"]))
    assert (img - torch.from_numpy(gold["img"])).abs().max() < 2e-5
    assert (ex["depth"] - torch.from_numpy(gold["depth"])).abs().max() < 2e-4
