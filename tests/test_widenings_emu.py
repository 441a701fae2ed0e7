"""SURVEY 8(f) kernels WITHOUT a GPU: the fused multi-tensor Adam step (adam.cu), the fused SSIM
forward / backward (ssim.cu) and the brute-force knn_points (knn.cu) run as the unchanged kernel
sources on the host SIMT emulator (tests/emu) and are compared with torch.optim.Adam, the SSIM
oracle (oracle/ssim_oracle.py) and an exhaustive search."""
import ctypes as C

import numpy as np
import pytest
import torch

import emu_lib
from emu_lib import ptr
from oracle import ssim_oracle


@pytest.fixture(scope="module")
def emu():
    return emu_lib.load()


def test_fused_adam_on_the_emulator_matches_torch_adam(emu):
    g = torch.Generator().manual_seed(0)
    shapes = [(700, 3), (333,), (129, 4), (0, 3), (50, 15, 3)]           # one empty tensor, ragged tails
    lrs = [1.6e-4, 5e-2, 1e-3, 1e-3, 1.25e-4]
    params = [torch.randn(*s, generator=g) for s in shapes]
    ref = [p.clone().requires_grad_(True) for p in params]
    opt = torch.optim.Adam([{"params": [p], "lr": lr} for p, lr in zip(ref, lrs)], eps=1e-15)
    mine = [np.ascontiguousarray(p.numpy().copy()) for p in params]
    m = [np.zeros_like(a) for a in mine]
    v = [np.zeros_like(a) for a in mine]
    n = len(shapes)
    for step in range(1, 4):
        grads = [torch.randn(*s, generator=g) for s in shapes]
        for p, gr in zip(ref, grads):
            p.grad = gr.clone()
        opt.step()
        gn = [np.ascontiguousarray(gr.numpy()) for gr in grads]
        arr = lambda xs: (C.c_void_p * n)(*[x.ctypes.data for x in xs])
        numels = (C.c_int64 * n)(*[a.size for a in mine])
        lr_c = (C.c_float * n)(*lrs)
        steps = (C.c_int64 * n)(*([step] * n))
        assert emu.emu_adam_step(n, arr(mine), arr(gn), arr(m), arr(v), numels, lr_c, steps, 0.9, 0.999, 1e-15) == 0
        for a, p in zip(mine, ref):
            if a.size:
                assert np.abs(a - p.detach().numpy()).max() <= 2e-6 * max(1.0, np.abs(a).max())


@pytest.mark.parametrize("B,Ch,H,W,hwc", [(1, 3, 40, 52, True), (2, 1, 27, 33, False)])
def test_fused_ssim_on_the_emulator_matches_the_oracle(emu, B, Ch, H, W, hwc):
    g = torch.Generator().manual_seed(1)
    if hwc:      # the rendered image is [H, W, 3]: read through its strides like tinysplat_b200.ssim does
        X = torch.rand(H, W, Ch, generator=g).permute(2, 0, 1)[None]
    else:
        X = torch.rand(B, Ch, H, W, generator=g)
    Y = (X + 0.1 * torch.randn(X.shape, generator=g)).clamp(0, 1).contiguous()
    Xr = X.detach().clone().requires_grad_(True)
    per = ssim_oracle.ssim_per_channel(Xr, Y)
    w_pc = torch.rand(per.shape, generator=g)
    (per * w_pc).sum().backward()

    win = np.ascontiguousarray(ssim_oracle.gaussian_window().numpy())
    Ho, Wo = H - 10, W - 10
    xs = (C.c_int64 * 4)(*X.stride())
    ys = (C.c_int64 * 4)(*Y.stride())
    Xn = X.numpy() if not hwc else np.ascontiguousarray(X[0].permute(1, 2, 0).numpy())   # storage order
    Yn = np.ascontiguousarray(Y.numpy())
    sums = np.zeros(B * Ch, np.float32)
    maps = [np.zeros((B, Ch, Ho, Wo), np.float32) for _ in range(3)]
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    assert emu.emu_ssim_fwd(B, Ch, H, W, ptr(Xn), xs, ptr(Yn), ys, ptr(win), C1, C2, ptr(sums), ptr(maps[0]),
                            ptr(maps[1]), ptr(maps[2])) == 0
    got = sums.reshape(B, Ch) / (Ho * Wo)
    assert np.abs(got - per.detach().numpy()).max() < 2e-5
    v_pc = np.ascontiguousarray(w_pc.numpy().reshape(-1).astype(np.float32))
    vX = np.zeros((B, Ch, H, W), np.float32)
    assert emu.emu_ssim_bwd(B, Ch, H, W, ptr(Xn), xs, ptr(Yn), ys, ptr(win), ptr(maps[0]), ptr(maps[1]), ptr(maps[2]),
                            ptr(v_pc), ptr(vX)) == 0
    want = Xr.grad.numpy()
    assert np.abs(vX - want).max() < 2e-4 * np.abs(want).max()


@pytest.mark.parametrize("P1,P2,K", [(300, 2500, 16), (33, 1500, 4), (130, 20, 16), (1, 1, 1)])
def test_knn_points_on_the_emulator_matches_exhaustive_search(emu, P1, P2, K):
    g = torch.Generator().manual_seed(P1 + P2)
    q = torch.randn(P1, 3, generator=g)
    r = torch.randn(P2, 3, generator=g)
    qn, rn = np.ascontiguousarray(q.numpy()), np.ascontiguousarray(r.numpy())
    d = np.zeros((P1, K), np.float32)
    idx = np.full((P1, K), -7, np.int64)
    assert emu.emu_knn_points(P1, P2, K, ptr(qn), ptr(rn), ptr(d), ptr(idx)) == 0
    full = ((q[:, None, :] - r[None, :, :]) ** 2).sum(-1)
    k = min(K, P2)
    wd, wi = full.topk(k, dim=1, largest=False)
    assert np.allclose(d[:, :k], wd.numpy(), rtol=1e-5, atol=1e-6)
    # indices may differ only where two references are equidistant
    same = idx[:, :k] == wi.numpy()
    assert same.mean() > 0.999
    assert np.allclose(full.gather(1, torch.from_numpy(idx[:, :k])).numpy(), wd.numpy(), rtol=1e-5, atol=1e-6)
    if k < K:
        assert (idx[:, k:] == 0).all() and (d[:, k:] == 0).all()


@pytest.mark.parametrize("n,blocks,u8", [(37 * 21 * 3, 3, False), (64 * 48 * 3, 5, True), (7, 1, True), (4096, 1, False)])
def test_fused_l1_loss_on_the_emulator(emu, n, blocks, u8):
    """csrc/loss.cu: mean |img - target| and its gradient in one pass, float32 or uint8 (/255) target,
    ragged tail, several blocks with the last-block fixed-order sum, the counter left at zero."""
    rng = np.random.default_rng(n)
    img = rng.random(n, dtype=np.float32)
    if u8:
        raw = rng.integers(0, 256, size=n, dtype=np.uint8)
        raw[: n // 3] = np.round(img[: n // 3] * 255).astype(np.uint8)       # some exact ties and near ties
        tgt = raw.astype(np.float32) / np.float32(255)
    else:
        raw = rng.random(n, dtype=np.float32)
        raw[::5] = img[::5]                                                   # exact zeros of the difference
        tgt = raw
    grad = np.full(n, 7.0, dtype=np.float32)
    work = np.zeros(emu.emu_l1_loss_work_floats(), dtype=np.float32)
    loss = np.zeros(1, dtype=np.float32)
    for _ in range(2):      # twice: the counter must have been reset
        assert emu.emu_l1_loss(n, emu_lib.ptr(img), emu_lib.ptr(raw), int(u8), 1.0 / n, 1.0 / n, emu_lib.ptr(grad),
                               emu_lib.ptr(work), emu_lib.ptr(loss), blocks) == 0
        d = img.astype(np.float64) - tgt.astype(np.float64)
        assert abs(loss[0] - np.abs(d).mean()) < 1e-6
        assert np.array_equal(grad, (np.sign(img - tgt) * np.float32(1.0 / n)).astype(np.float32))
