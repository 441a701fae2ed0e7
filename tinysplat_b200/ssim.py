"""Fused SSIM for the training loss.

Stand-in for `pytorch_msssim.SSIM` with the constructor arguments the reference uses,
`SSIM(data_range=1.0, size_average=True, channel=3)` [REF tinysplat/splatting/model_gaussian.py:13,57],
called as `model.ssim(rendered[1,3,H,W], gt[1,3,H,W])` [REF scripts/train.py:60-62]:

    from tinysplat_b200.ssim import SSIM

Forward and backward are one kernel each; the [H,W,3] rendered image is read through its strides
(no permute/contiguous copy).  Gradients flow to the first argument only (the ground truth is a
constant).  No CPU path."""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import torch
from torch import Tensor, nn
from torch.autograd import Function

from . import _lib

WIN = 11


def _window(sigma: float):
    coords = torch.arange(WIN, dtype=torch.float32) - WIN // 2
    g = torch.exp(-(coords ** 2) / (2 * sigma ** 2))
    g = g / g.sum()
    return (C.c_float * WIN)(*g.tolist())


def _strides(t: Tensor):
    return (C.c_int64 * 4)(*t.stride())


class _FusedSSIM(Function):
    @staticmethod
    def forward(ctx, X: Tensor, Y: Tensor, C1: float, C2: float, sigma: float):
        _lib.require_cuda(X, Y)
        if X.dim() != 4 or X.shape != Y.shape:
            raise ValueError("SSIM expects two [B, C, H, W] tensors of the same shape")
        B, Ch, H, W = X.shape
        if H <= WIN - 1 or W <= WIN - 1:
            raise ValueError("image smaller than the 11x11 SSIM window")
        Xd = X.detach() if X.dtype == torch.float32 else X.detach().float()
        Yd = Y.detach() if Y.dtype == torch.float32 else Y.detach().float()
        dev = X.device
        win = _window(sigma)
        need_grad = ctx.needs_input_grad[0]
        Ho, Wo = H - (WIN - 1), W - (WIN - 1)
        sums = torch.empty(B * Ch, device=dev, dtype=torch.float32)
        maps = [torch.empty(B, Ch, Ho, Wo, device=dev, dtype=torch.float32) for _ in range(3)] if need_grad \
            else [None, None, None]
        _lib.call("ts_ssim_fwd", B, Ch, H, W, _lib.ptr(Xd), _strides(Xd), _lib.ptr(Yd), _strides(Yd), win,
                  float(C1), float(C2), _lib.ptr(sums), _lib.ptr(maps[0]), _lib.ptr(maps[1]), _lib.ptr(maps[2]),
                  _lib.stream_ptr(dev))
        if need_grad:
            ctx.save_for_backward(Xd, Yd, *maps)
        ctx.meta = (B, Ch, H, W, sigma)
        return sums.view(B, Ch) / float(Ho * Wo)

    @staticmethod
    def backward(ctx, v_pc: Tensor):
        if ctx.needs_input_grad[1]:
            raise NotImplementedError("fused SSIM differentiates with respect to its first argument only")
        Xd, Yd, dmu, de11, de12 = ctx.saved_tensors
        B, Ch, H, W, sigma = ctx.meta
        v = _lib.f32c(v_pc)
        v_X = torch.empty(B, Ch, H, W, device=Xd.device, dtype=torch.float32)
        _lib.call("ts_ssim_bwd", B, Ch, H, W, _lib.ptr(Xd), _strides(Xd), _lib.ptr(Yd), _strides(Yd),
                  _window(sigma), _lib.ptr(dmu), _lib.ptr(de11), _lib.ptr(de12), _lib.ptr(v), _lib.ptr(v_X),
                  _lib.stream_ptr(Xd.device))
        return v_X, None, None, None, None


def ssim(X: Tensor, Y: Tensor, data_range: float = 255, size_average: bool = True, win_size: int = 11,
         win_sigma: float = 1.5, K: Sequence[float] = (0.01, 0.03), nonnegative_ssim: bool = False) -> Tensor:
    if win_size != WIN:
        raise NotImplementedError("fused SSIM supports the 11-tap window only")
    C1, C2 = (K[0] * data_range) ** 2, (K[1] * data_range) ** 2
    per_channel = _FusedSSIM.apply(X, Y, C1, C2, float(win_sigma))
    if nonnegative_ssim:
        per_channel = torch.relu(per_channel)
    return per_channel.mean() if size_average else per_channel.mean(1)


class SSIM(nn.Module):
    """Same constructor / call as pytorch_msssim.SSIM for the 2-D, 11-tap case."""

    def __init__(self, data_range: float = 255, size_average: bool = True, win_size: int = 11,
                 win_sigma: float = 1.5, channel: int = 3, spatial_dims: int = 2,
                 K: Sequence[float] = (0.01, 0.03), nonnegative_ssim: bool = False):
        super().__init__()
        if spatial_dims != 2:
            raise NotImplementedError("fused SSIM supports 2-D images only")
        self.kw = dict(data_range=data_range, size_average=size_average, win_size=win_size,
                       win_sigma=win_sigma, K=tuple(K), nonnegative_ssim=nonnegative_ssim)
        self.channel = channel

    def forward(self, X: Tensor, Y: Tensor) -> Tensor:
        return ssim(X, Y, **self.kw)
