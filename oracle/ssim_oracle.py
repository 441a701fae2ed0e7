"""CPU restatement of the SSIM the reference trains with — TEST INFRASTRUCTURE ONLY.

The reference builds `pytorch_msssim.SSIM(data_range=1.0, size_average=True, channel=3)`
[REF tinysplat/splatting/model_gaussian.py:13,57] and calls it on [1,3,H,W] views of the
rendered and ground-truth images [REF scripts/train.py:60-62].  pytorch_msssim is a third-party
dependency that is absent here (not vendored, not pinned, not installable): PARITY UNPINNED.
This file restates its published algorithm (Wang et al. 2004 SSIM; 11-tap Gaussian window,
sigma 1.5, separable, 'valid' filtering, K = (0.01, 0.03)) and is pinned by closed forms and
gradcheck in tests/test_ssim.py."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def gaussian_window(size: int = 11, sigma: float = 1.5, dtype=torch.float32) -> torch.Tensor:
    coords = torch.arange(size, dtype=dtype) - size // 2
    g = torch.exp(-(coords ** 2) / (2 * sigma ** 2))
    return g / g.sum()


def _filter(x: torch.Tensor, win: torch.Tensor) -> torch.Tensor:
    """Separable 'valid' Gaussian filtering of [B,C,H,W]."""
    C = x.shape[1]
    w = win.to(x.dtype)
    x = F.conv2d(x, w.view(1, 1, -1, 1).repeat(C, 1, 1, 1), groups=C)
    x = F.conv2d(x, w.view(1, 1, 1, -1).repeat(C, 1, 1, 1), groups=C)
    return x


def ssim_per_channel(X: torch.Tensor, Y: torch.Tensor, data_range: float = 1.0, win_size: int = 11,
                     win_sigma: float = 1.5, K=(0.01, 0.03)) -> torch.Tensor:
    """[B,C] mean SSIM per channel."""
    win = gaussian_window(win_size, win_sigma, torch.float32)
    C1, C2 = (K[0] * data_range) ** 2, (K[1] * data_range) ** 2
    mu1, mu2 = _filter(X, win), _filter(Y, win)
    s1 = _filter(X * X, win) - mu1 * mu1
    s2 = _filter(Y * Y, win) - mu2 * mu2
    s12 = _filter(X * Y, win) - mu1 * mu2
    cs = (2 * s12 + C2) / (s1 + s2 + C2)
    m = ((2 * mu1 * mu2 + C1) / (mu1 * mu1 + mu2 * mu2 + C1)) * cs
    return m.flatten(2).mean(-1)


def ssim(X, Y, data_range=1.0, size_average=True, nonnegative_ssim=False, **kw):
    s = ssim_per_channel(X, Y, data_range, **kw)
    if nonnegative_ssim:
        s = torch.relu(s)
    return s.mean() if size_average else s.mean(1)
