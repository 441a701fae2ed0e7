"""Fused L1 image loss (the caller side of the hot path).

The reference's step computes `(rendered_image - ground_truth_image).abs().mean()` with the ground
truth uploaded as float32 every step [REF scripts/train.py:58-59; tinysplat/scene.py:27-31,130-132]:
three elementwise kernels forward, three backward, 100 bytes of HBM traffic per element.  Here forward
and gradient are one pass (csrc/loss.cu), and the ground truth may stay the uint8 image it was loaded
as (`value = u8 / 255` is formed in the kernel, identical to torch's fp32 division): a quarter of the
host->device bytes per step.

    from tinysplat_b200.loss import l1_loss
    loss = l1_loss(rendered, gt)            # gt: float32 in [0, 1] or uint8, same shape as rendered

Gradients flow to the first argument only.  Deterministic (no float atomics).  No CPU path."""
from __future__ import annotations

import torch
from torch import Tensor
from torch.autograd import Function

from . import _lib

_work = {}


def _work_buffer(dev) -> Tensor:
    key = (dev.index if dev.index is not None else torch.cuda.current_device(), torch.cuda.current_stream(dev).cuda_stream)
    w = _work.get(key)
    if w is None:
        w = _work[key] = torch.zeros(_lib.load().ts_l1_loss_work_floats(), device=dev, dtype=torch.float32)
    return w


class _L1Loss(Function):
    @staticmethod
    def forward(ctx, img: Tensor, target: Tensor):
        _lib.require_cuda(img, target)
        if img.shape != target.shape:
            raise ValueError(f"l1_loss: shapes differ, {tuple(img.shape)} vs {tuple(target.shape)}")
        if target.dtype not in (torch.float32, torch.uint8):
            raise TypeError("l1_loss: the ground truth must be float32 or uint8")
        x = _lib.f32c(img.detach())
        t = target.detach().contiguous()
        n = x.numel()
        if n == 0:
            raise ValueError("l1_loss of an empty image")
        dev = x.device
        need_grad = ctx.needs_input_grad[0]
        grad = torch.empty_like(x) if need_grad else None
        loss = torch.empty((), device=dev, dtype=torch.float32)
        _lib.call("ts_l1_loss", n, _lib.ptr(x), _lib.ptr(t), 1 if t.dtype == torch.uint8 else 0, 1.0 / n, 1.0 / n,
                  _lib.ptr(grad), _lib.ptr(_work_buffer(dev)), _lib.ptr(loss), _lib.stream_ptr(dev))
        if need_grad:
            ctx.save_for_backward(grad)
        ctx.in_shape = img.shape
        return loss

    @staticmethod
    def backward(ctx, v):
        (grad,) = ctx.saved_tensors
        return (grad * v).view(ctx.in_shape), None


def l1_loss(rendered: Tensor, ground_truth: Tensor) -> Tensor:
    """mean |rendered - ground_truth| ; ground_truth float32, or uint8 meaning value / 255."""
    return _L1Loss.apply(rendered, ground_truth)
