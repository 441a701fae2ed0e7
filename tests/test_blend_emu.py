"""Kernel-logic parity WITHOUT a GPU: the blend kernels (tinysplat_b200/csrc/blend.cu: forward and
first-generation backward; blend_group.cu: grouped backward) are compiled as host code on the
fiber SIMT emulator in tests/emu (threadIdx, __shared__, warp votes / shuffles, barriers, cp.async
as memcpy) and compared with the oracle on small scenes.  What this covers: staging, sub-block /
exact row masks, the list walks, T-termination, n_contrib, the shared-memory group reduction and
the 8-lane transpose-reduce, the packed-gradient layout.  What it cannot cover (memory model,
real async copies, occupancy) is left to the `-m gpu` tests."""
import ctypes as C
import subprocess

import numpy as np
import pytest
import torch

import oracle
from oracle import gsplat_oracle as go
from tinysplat_b200 import synthetic

import emu_lib

LOG2E = 1.4426950408889634


@pytest.fixture(scope="module")
def emu():
    return emu_lib.load()


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _scene(n, W, H, seed, radius_px, channels):
    cam = synthetic.make_camera(W, H)
    sc = synthetic.make_scene(n, W, H, seed=seed, mean_radius_px=radius_px)
    tb = ((W + 15) // 16, (H + 15) // 16, 1)
    with torch.no_grad():
        xys, depths, radii, conics, ntiles, _ = oracle.project_gaussians(
            sc["means"], sc["scales"].exp(), 1.0, torch.nn.functional.normalize(sc["quats"], dim=-1),
            cam.view_matrix[:3], cam.proj_matrix @ cam.view_matrix, cam.f_x, cam.f_y, W / 2, H / 2, H, W, tb)
    g = torch.Generator().manual_seed(seed + 100)
    colors = torch.rand(n, channels, generator=g)
    opac = torch.sigmoid(sc["opacities"])
    return xys, depths, radii, conics, ntiles, colors, opac, tb


def _pack(xys, conics, opac, colors, cull):
    """The packed raster record of ts_common.cuh (what ts_bin_count / ts_project_fwd write)."""
    n, ch = colors.shape
    a, b, c = conics[:, 0].double(), conics[:, 1].double(), conics[:, 2].double()
    op = opac.reshape(-1).double()
    rec = np.zeros((n, 12), dtype=np.float32)
    rec[:, 0:2] = xys.numpy()
    if cull:
        det = a * c - b * b
        two_tau = 2.0 * (torch.log(255.0 * op) + 0.01)
        hx = torch.sqrt(two_tau * c / det) * 1.001 + 0.01
        hy = torch.sqrt(two_tau * a / det) * 1.001 + 0.01
        bad = ~(op * 255.0 >= 1.0)
        hx[bad], hy[bad] = -1e30, -1e30
        nogood = ~((det > 0) & (a > 0) & (c > 0)) & ~bad
        hx[nogood], hy[nogood] = 1e30, 1e30
        rec[:, 2], rec[:, 3] = hx.float().numpy(), hy.float().numpy()
    else:
        rec[:, 2:4] = 1e30
    cf = conics.float()
    rec[:, 4] = (np.float32(0.5 * LOG2E) * cf[:, 0]).numpy()
    rec[:, 5] = (np.float32(LOG2E) * cf[:, 1]).numpy()
    rec[:, 6] = (np.float32(0.5 * LOG2E) * cf[:, 2]).numpy()
    rec[:, 7] = opac.reshape(-1).float().numpy()
    rec[:, 8:8 + ch] = colors.numpy()
    return rec


def _lists(xys, depths, radii, tb):
    tile, gid = go.bin_and_sort(xys, depths, radii, tb)
    T = tb[0] * tb[1]
    offsets = np.zeros(T + 1, dtype=np.int32)
    offsets[1:] = np.cumsum(np.bincount(tile.numpy(), minlength=T))
    ids = np.ascontiguousarray(gid.numpy().astype(np.int32))
    if ids.size == 0:
        ids = np.zeros(1, dtype=np.int32)
    return offsets, ids


def _unpack(grads, conics, ch):
    g = torch.from_numpy(grads).double()
    a, b, c = conics[:, 0].double(), conics[:, 1].double(), conics[:, 2].double()
    v_xys = torch.stack([a * g[:, 0] + b * g[:, 1], b * g[:, 0] + c * g[:, 1]], -1)
    v_conics = torch.stack([0.5 * g[:, 2], g[:, 3], 0.5 * g[:, 4]], -1)
    return v_xys, v_conics, g[:, 8:8 + ch], g[:, 5]


def _rel(x, ref):
    return (x - ref).abs().max().item() / max(ref.abs().max().item(), 1e-12)


CASES = [
    # n, W, H, radius, channels, cull
    (300, 48, 40, 5.0, 3, True),      # ragged image (partial tiles right and bottom)
    (300, 48, 40, 5.0, 3, False),     # culling disabled: every bit set, bounding-free path
    (200, 64, 32, 12.0, 4, True),     # big splats, > 128 candidates per tile (several batches)
    (2500, 32, 32, 4.0, 1, True),     # deep tiles: T-termination, many batches
    (64, 16, 16, 3.0, 2, True),       # a single tile
]


@pytest.mark.parametrize("n,W,H,radius,ch,cull", CASES)
def test_blend_kernels_match_the_oracle_on_the_emulator(emu, n, W, H, radius, ch, cull):
    xys, depths, radii, conics, ntiles, colors, opac, tb = _scene(n, W, H, 3, radius, ch)
    bg = torch.linspace(0.1, 0.7, ch)
    rec = _pack(xys, conics, opac, colors, cull)
    offsets, ids = _lists(xys, depths, radii, tb)
    out = np.full((H, W, ch), -7.0, dtype=np.float32)
    final_T = np.full((H, W), -7.0, dtype=np.float32)
    ncon = np.full((H, W), -7, dtype=np.int32)
    bgn = np.ascontiguousarray(bg.numpy())
    assert emu.emu_blend_fwd(ch, H, W, tb[0], tb[1], _ptr(offsets), _ptr(ids), _ptr(rec), _ptr(bgn), _ptr(out),
                             None, _ptr(final_T), _ptr(ncon), 0) == 0

    # oracle forward + autograd
    x = xys.clone().requires_grad_(True)
    cn = conics.clone().requires_grad_(True)
    co = colors.clone().requires_grad_(True)
    o = opac.clone().requires_grad_(True)
    img, alpha, aux = oracle.rasterize_gaussians(x, depths, radii, cn, ntiles, co, o, H, W, bg, return_aux=True)
    assert np.abs(out - img.detach().numpy()).max() < 2e-5
    assert np.abs((1.0 - final_T) - alpha.detach().numpy()).max() < 2e-5
    assert np.array_equal(ncon, aux["n_contrib"].numpy().astype(np.int32))
    # launch order: longest tile lists first (ts_bin_tile_order) must give the same bits as raster order
    T = tb[0] * tb[1]
    order = np.full(T, -1, dtype=np.int32)
    assert emu.emu_bin_tile_order(T, _ptr(offsets), _ptr(order)) == 0
    assert np.array_equal(np.sort(order), np.arange(T))                      # a permutation
    counts = np.diff(offsets)[order]
    assert (np.diff(np.minimum(counts, 2047) // 8) <= 0).all()               # by descending length class
    out1, T1, nc1 = np.full_like(out, -7.0), np.full_like(final_T, -7.0), np.full_like(ncon, -7)
    emu.emu_set_tile_order(_ptr(order))
    try:
        assert emu.emu_blend_fwd(ch, H, W, tb[0], tb[1], _ptr(offsets), _ptr(ids), _ptr(rec), _ptr(bgn), _ptr(out1),
                                 None, _ptr(T1), _ptr(nc1), 0) == 0
    finally:
        emu.emu_set_tile_order(None)
    assert np.array_equal(out1, out) and np.array_equal(T1, final_T) and np.array_equal(nc1, ncon)

    g = torch.Generator().manual_seed(9)
    v_img = torch.rand(H, W, ch, generator=g)
    v_alpha = torch.rand(H, W, generator=g)
    (img * v_img).sum().add((alpha * v_alpha).sum()).backward()
    vi = np.ascontiguousarray(v_img.numpy())
    va = np.ascontiguousarray(v_alpha.numpy())
    for direct in (0, 1):     # first-generation backward / grouped backward
        grads = np.zeros((n, 12), dtype=np.float32)
        assert emu.emu_blend_bwd(n, ch, H, W, tb[0], tb[1], _ptr(offsets), _ptr(ids), _ptr(rec), _ptr(bgn),
                                 _ptr(final_T), _ptr(ncon), _ptr(vi), None, 0, _ptr(va), _ptr(grads), direct) == 0
        v_xys, v_conics, v_colors, v_opac = _unpack(grads, conics, ch)
        assert _rel(v_xys, x.grad.double()) < 1e-4
        assert _rel(v_conics, cn.grad.double()) < 1e-4
        assert _rel(v_colors, co.grad.double()) < 1e-4
        assert _rel(v_opac, o.grad.reshape(-1).double()) < 1e-4


def test_split_rgb_depth_pass_with_clamp_on_the_emulator(emu):
    """The fused pipeline's 4-channel pass: RGB image + separate depth map, clamp(rgb, max=1)
    folded in (clamped channels get no gradient), with and without a depth cotangent."""
    n, W, H = 400, 48, 32
    xys, depths, radii, conics, ntiles, colors, opac, tb = _scene(n, W, H, 5, 6.0, 4)
    colors[:, :3] *= 2.5                      # make the clamp bite
    colors[:, 3] = depths
    bg = torch.tensor([0.2, 0.5, 0.8, 0.0])
    rec = _pack(xys, conics, opac, colors, True)
    offsets, ids = _lists(xys, depths, radii, tb)
    out = np.zeros((H, W, 3), dtype=np.float32)
    out3 = np.zeros((H, W), dtype=np.float32)
    final_T = np.zeros((H, W), dtype=np.float32)
    ncon = np.zeros((H, W), dtype=np.int32)
    bgn = np.ascontiguousarray(bg.numpy())
    assert emu.emu_blend_fwd(4, H, W, tb[0], tb[1], _ptr(offsets), _ptr(ids), _ptr(rec), _ptr(bgn), _ptr(out),
                             _ptr(out3), _ptr(final_T), _ptr(ncon), 1) == 0
    g = torch.Generator().manual_seed(2)
    v_img = torch.rand(H, W, 3, generator=g)
    v_dep = torch.rand(H, W, generator=g)
    for with_depth, direct in ((False, 0), (True, 0), (False, 1), (True, 1)):
        x = xys.clone().requires_grad_(True)
        cn = conics.clone().requires_grad_(True)
        co = colors.clone().requires_grad_(True)
        o = opac.clone().requires_grad_(True)
        img, _ = oracle.rasterize_gaussians(x, depths, radii, cn, ntiles, co, o, H, W, bg)
        rgb = torch.clamp(img[..., :3], max=1.0)
        assert np.abs(out - rgb.detach().numpy()).max() < 2e-5
        assert np.abs(out3 - img[..., 3].detach().numpy()).max() < 2e-4
        assert (rgb.detach() == 1.0).any()
        loss = (rgb * v_img).sum()
        if with_depth:
            loss = loss + (img[..., 3] * v_dep).sum()
        loss.backward()
        grads = np.zeros((n, 12), dtype=np.float32)
        vi = np.ascontiguousarray(v_img.numpy())
        vd = np.ascontiguousarray(v_dep.numpy()) if with_depth else None
        assert emu.emu_blend_bwd(n, 4, H, W, tb[0], tb[1], _ptr(offsets), _ptr(ids), _ptr(rec), _ptr(bgn),
                                 _ptr(final_T), _ptr(ncon), _ptr(vi), _ptr(vd), 1, None, _ptr(grads), direct) == 0
        v_xys, v_conics, v_colors, v_opac = _unpack(grads, conics, 4)
        assert _rel(v_xys, x.grad.double()) < 1e-4
        assert _rel(v_conics, cn.grad.double()) < 1e-4
        assert _rel(v_colors, co.grad.double()) < 1e-4
        assert _rel(v_opac, o.grad.reshape(-1).double()) < 1e-4


def test_emulator_detects_a_lane_that_skips_a_collective(emu, tmp_path):
    """The emulator's own safety net: a kernel in which one lane skips a warp vote must be
    reported as a deadlock, not hang the test run."""
    src = tmp_path / "dead.cpp"
    src.write_text('#include "ts_emu.h"\n'
                   'extern "C" int run() { return ts_emu::launch(dim3(1, 1), 32, []() {\n'
                   '  if (threadIdx.x != 5) __ballot_sync(0xffffffffu, 1); }); }\n')
    so = tmp_path / "dead.so"
    subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-I" + emu_lib.EMU, "-o", str(so), str(src)], check=True)
    assert C.CDLL(str(so)).run() == -1


def test_backward_zero_opacity_without_culling_stays_finite(emu):
    """cull_mode = 0 keeps every (Gaussian, tile) pair, including Gaussians whose opacity is
    exactly 0: v_opacity is formed as -s0 / opacity in the grouped kernel and must not turn a
    0 * inf into a NaN that a neighbouring group's reduction then spreads."""
    n, W, H, ch = 200, 32, 32, 3
    xys, depths, radii, conics, ntiles, colors, opac, tb = _scene(n, W, H, 8, 6.0, ch)
    opac[::3] = 0.0
    bg = torch.zeros(ch)
    rec = _pack(xys, conics, opac, colors, False)
    offsets, ids = _lists(xys, depths, radii, tb)
    out = np.zeros((H, W, ch), dtype=np.float32)
    final_T = np.zeros((H, W), dtype=np.float32)
    ncon = np.zeros((H, W), dtype=np.int32)
    bgn = np.ascontiguousarray(bg.numpy())
    assert emu.emu_blend_fwd(ch, H, W, tb[0], tb[1], _ptr(offsets), _ptr(ids), _ptr(rec), _ptr(bgn), _ptr(out),
                             None, _ptr(final_T), _ptr(ncon), 0) == 0
    vi = np.ones((H, W, ch), dtype=np.float32)
    for direct in (0, 1):
        grads = np.zeros((n, 12), dtype=np.float32)
        assert emu.emu_blend_bwd(n, ch, H, W, tb[0], tb[1], _ptr(offsets), _ptr(ids), _ptr(rec), _ptr(bgn),
                                 _ptr(final_T), _ptr(ncon), _ptr(vi), None, 0, None, _ptr(grads), direct) == 0
        assert np.isfinite(grads).all()
        assert np.abs(grads[::3]).max() == 0.0
        assert np.abs(grads).max() > 0.0
