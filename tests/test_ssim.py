"""SSIM: CPU pins of the oracle (parity unpinned upstream: pytorch_msssim is absent)
and GPU parity of the fused kernels against it."""
import pytest
import torch

from oracle import ssim_oracle as so

DEV = "cuda:0"


def test_oracle_window_and_identity():
    w = so.gaussian_window()
    assert w.shape == (11,) and abs(float(w.sum()) - 1) < 1e-6 and torch.allclose(w, w.flip(0))
    x = torch.rand(2, 3, 40, 33, dtype=torch.float64)
    assert torch.allclose(so.ssim(x, x), torch.tensor(1.0, dtype=torch.float64), atol=1e-12)
    y = torch.rand(2, 3, 40, 33, dtype=torch.float64)
    assert torch.allclose(so.ssim(x, y), so.ssim(y, x), atol=1e-12)          # symmetric
    assert so.ssim_per_channel(x, y).shape == (2, 3)
    assert float(so.ssim(x, y)) < 0.2                                        # independent noise


def test_oracle_constant_images_closed_form():
    # constant images: sigma terms vanish, SSIM = (2ab + C1)/(a^2 + b^2 + C1)
    a, b = 0.3, 0.8
    x = torch.full((1, 1, 20, 20), a, dtype=torch.float64)
    y = torch.full((1, 1, 20, 20), b, dtype=torch.float64)
    C1 = 0.01 ** 2
    assert abs(float(so.ssim(x, y)) - (2 * a * b + C1) / (a * a + b * b + C1)) < 5e-5   # fp32 window weights sum to 1 +- 6e-8; C2 = 9e-4 amplifies it


def test_oracle_gradcheck():
    x = torch.rand(1, 2, 14, 13, dtype=torch.float64, requires_grad=True)
    y = torch.rand(1, 2, 14, 13, dtype=torch.float64)
    assert torch.autograd.gradcheck(lambda t: so.ssim_per_channel(t, y), (x,), atol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("B,H,W", [(1, 64, 80), (2, 37, 53), (1, 11, 11), (1, 128, 128)])
def test_fused_ssim_matches_oracle(lib, B, H, W):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from tinysplat_b200.ssim import SSIM
    g = torch.Generator().manual_seed(H * W)
    img = torch.rand(B, H, W, 3, generator=g)                       # HWC, as the rasterizer returns it
    gt = (img + 0.2 * torch.randn(B, H, W, 3, generator=g)).clamp(0, 1)
    x_ref = img.double().permute(0, 3, 1, 2).clone().requires_grad_(True)
    want = so.ssim(x_ref, gt.double().permute(0, 3, 1, 2), data_range=1.0)
    (1 - want).backward()
    leaf = img.to(DEV).requires_grad_(True)
    mod = SSIM(data_range=1.0, size_average=True, channel=3)
    got = mod(leaf.permute(0, 3, 1, 2), gt.to(DEV).permute(0, 3, 1, 2))          # strided views, no copies
    (1 - got).backward()
    assert abs(float(got) - float(want)) < 2e-6
    ref_grad = x_ref.grad.permute(0, 2, 3, 1)
    err = (leaf.grad.cpu().double() - ref_grad).abs().max().item() / ref_grad.abs().max().item()
    assert err < 1e-4, err
    with torch.no_grad():
        assert abs(float(mod(leaf.permute(0, 3, 1, 2), gt.to(DEV).permute(0, 3, 1, 2))) - float(want)) < 2e-6
