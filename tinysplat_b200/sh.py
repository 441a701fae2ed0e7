"""gsplat.sh surface: spherical_harmonics, num_sh_bases, deg_from_sh
[REF tinysplat/splatting/rasterize.py:3,38,76; tinysplat/splatting/model_gaussian.py:14,71,106]."""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor
from torch.autograd import Function

from . import _lib


def num_sh_bases(degree: int) -> int:
    """Number of real SH bases up to `degree` (0..4): (degree+1)^2."""
    if degree < 0 or degree > 4:
        raise ValueError(f"SH degree must be in 0..4, got {degree}")
    return (degree + 1) ** 2


def deg_from_sh(num_bases: int) -> int:
    """Inverse of num_sh_bases."""
    for deg in range(5):
        if (deg + 1) ** 2 == num_bases:
            return deg
    raise ValueError(f"Invalid number of SH bases: {num_bases}")


class _SphericalHarmonics(Function):
    """colors[N,3] = sum_k Y_k(dir) coeffs[N,k,:].  `rest` (optional) lets the caller keep
    the DC band and the higher bands in two tensors, as GaussianModel stores them
    [REF model_gaussian.py:86-87], without the per-step torch.cat [REF rasterize.py:80]."""

    @staticmethod
    def forward(ctx, degree: int, dirs: Tensor, coeffs: Tensor, rest: Optional[Tensor]):
        _lib.require_cuda(dirs, coeffs)
        lib = _lib.load()
        N = coeffs.shape[0]
        K = coeffs.shape[-2] + (rest.shape[-2] if rest is not None else 0)
        if K < num_sh_bases(degree):
            raise ValueError("coeffs has fewer SH bases than degrees_to_use needs")
        dirs_c = _lib.f32c(dirs.detach())
        co = _lib.f32c(coeffs.detach())
        re = _lib.f32c(rest.detach()) if rest is not None else None
        colors = torch.empty(N, 3, device=coeffs.device, dtype=torch.float32)
        _lib.call("ts_sh_fwd", N, degree, K, _lib.ptr(dirs_c), None, _lib.ptr(co), _lib.ptr(re),
                  _lib.ptr(colors), 3, None, None, 0, _lib.stream_ptr(coeffs.device))
        ctx.save_for_backward(dirs_c)
        ctx.meta = (degree, K, N, rest is not None)
        return colors

    @staticmethod
    def backward(ctx, v_colors: Tensor):
        (dirs_c,) = ctx.saved_tensors
        degree, K, N, split = ctx.meta
        lib = _lib.load()
        v_colors = _lib.f32c(v_colors)
        dev = v_colors.device
        if split:
            v_dc = torch.empty(N, 1, 3, device=dev, dtype=torch.float32)
            v_rest = torch.empty(N, K - 1, 3, device=dev, dtype=torch.float32)
        else:
            v_dc = torch.empty(N, K, 3, device=dev, dtype=torch.float32)
            v_rest = None
        _lib.call("ts_sh_bwd", N, degree, K, _lib.ptr(dirs_c), None, _lib.ptr(v_colors), 3, None,
                  _lib.ptr(v_dc), _lib.ptr(v_rest), 0, _lib.stream_ptr(dev))
        # view directions get no gradient (recorded in DESIGN.md)
        return None, None, v_dc, v_rest


def spherical_harmonics(degrees_to_use: int, viewdirs: Tensor, coeffs: Tensor) -> Tensor:
    """Drop-in for gsplat.sh.spherical_harmonics(degree, dirs[N,3], coeffs[N,K,3]) -> [N,3]
    [REF rasterize.py:38,81].  The caller adds 0.5 and clamps [REF rasterize.py:39]."""
    if coeffs.dim() != 3 or coeffs.shape[-1] != 3:
        raise ValueError("coeffs must be [N, K, 3]")
    return _SphericalHarmonics.apply(int(degrees_to_use), viewdirs, coeffs, None)


def spherical_harmonics_split(degrees_to_use: int, viewdirs: Tensor, coeffs_dc: Tensor,
                              coeffs_rest: Tensor) -> Tensor:
    """Same op with the DC band [N,3] and the rest [N,K-1,3] passed separately."""
    return _SphericalHarmonics.apply(int(degrees_to_use), viewdirs, coeffs_dc.unsqueeze(1)
                                     if coeffs_dc.dim() == 2 else coeffs_dc, coeffs_rest)
