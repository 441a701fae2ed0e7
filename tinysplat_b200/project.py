"""gsplat.project_gaussians drop-in  [REF tinysplat/splatting/rasterize.py:4,32,64-73]."""
from __future__ import annotations

from typing import Sequence, Tuple

import torch
from torch import Tensor
from torch.autograd import Function

from . import _lib


class _ProjectGaussians(Function):
    @staticmethod
    def forward(ctx, means3d, scales, glob_scale, quats, viewmat, projmat, fx, fy, cx, cy,
                img_height, img_width, tile_bounds, clip_thresh):
        _lib.require_cuda(means3d, scales, quats)
        lib = _lib.load()
        dev = means3d.device
        N = means3d.shape[0]
        means_c = _lib.f32c(means3d.detach())
        scales_c = _lib.f32c(scales.detach())
        quats_c = _lib.f32c(quats.detach())
        view_c = _lib.f32c(viewmat.detach().to(dev))
        proj_c = _lib.f32c(projmat.detach().to(dev))
        if view_c.numel() < 12 or proj_c.numel() != 16:
            raise ValueError("viewmat must be [3,4] or [4,4] and projmat [4,4]")
        f32 = dict(device=dev, dtype=torch.float32)
        i32 = dict(device=dev, dtype=torch.int32)
        xys = torch.empty(N, 2, **f32)
        depths = torch.empty(N, **f32)
        radii = torch.empty(N, **i32)
        conics = torch.empty(N, 3, **f32)
        num_tiles_hit = torch.empty(N, **i32)
        cov3d = torch.empty(N, 6, **f32)
        args = (float(glob_scale), float(fx), float(fy), float(cx), float(cy), int(img_height),
                int(img_width))
        _lib.call("ts_project_fwd", N, _lib.ptr(means_c), _lib.ptr(scales_c), args[0], _lib.ptr(quats_c),
            _lib.ptr(view_c), _lib.ptr(proj_c), args[1], args[2], args[3], args[4], args[5], args[6],
            int(tile_bounds[0]), int(tile_bounds[1]), float(clip_thresh), 0,
            _lib.ptr(xys), _lib.ptr(depths), _lib.ptr(radii), _lib.ptr(conics),
            _lib.ptr(num_tiles_hit), _lib.ptr(cov3d), None, 0, None, None, _lib.stream_ptr(dev))
        ctx.save_for_backward(means_c, scales_c, quats_c, view_c, proj_c, radii)
        ctx.args = args
        ctx.mark_non_differentiable(radii, num_tiles_hit, cov3d)
        return xys, depths, radii, conics, num_tiles_hit, cov3d

    @staticmethod
    def backward(ctx, v_xys, v_depths, v_radii, v_conics, v_num_tiles_hit, v_cov3d):
        means_c, scales_c, quats_c, view_c, proj_c, radii = ctx.saved_tensors
        gs, fx, fy, cx, cy, H, W = ctx.args
        lib = _lib.load()
        dev = means_c.device
        N = means_c.shape[0]
        v_xys = _lib.f32c(v_xys) if v_xys is not None else None
        v_depths = _lib.f32c(v_depths) if v_depths is not None else None
        v_conics = _lib.f32c(v_conics) if v_conics is not None else None
        v_means = torch.empty(N, 3, device=dev, dtype=torch.float32)
        v_scales = torch.empty(N, 3, device=dev, dtype=torch.float32)
        v_quats = torch.empty(N, 4, device=dev, dtype=torch.float32)
        _lib.call("ts_project_bwd", N, _lib.ptr(means_c), _lib.ptr(scales_c), gs, _lib.ptr(quats_c), _lib.ptr(view_c),
            _lib.ptr(proj_c), fx, fy, cx, cy, H, W, 0, _lib.ptr(radii), _lib.ptr(v_xys),
            _lib.ptr(v_depths), _lib.ptr(v_conics), None, None, _lib.ptr(v_means), _lib.ptr(v_scales),
            _lib.ptr(v_quats), None, None, _lib.stream_ptr(dev))
        return (v_means, v_scales, None, v_quats) + (None,) * 10


def project_gaussians(means3d: Tensor, scales: Tensor, glob_scale: float, quats: Tensor,
                      viewmat: Tensor, projmat: Tensor, fx: float, fy: float, cx: float,
                      cy: float, img_height: int, img_width: int,
                      tile_bounds: Sequence[int], clip_thresh: float = 0.01
                      ) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor, Tensor]:
    """EWA projection of N Gaussians.  Positional signature exactly as tinysplat calls it
    [REF rasterize.py:73] (note img_height BEFORE img_width).  Returns the 6-tuple
    (xys[N,2], depths[N], radii[N] int32, conics[N,3], num_tiles_hit[N] int32, cov3d[N,6])
    [REF rasterize.py:32].  quats are (w,x,y,z) and must already be unit length (the caller
    normalises in torch [REF rasterize.py:73])."""
    if means3d.dim() != 2 or means3d.shape[1] != 3:
        raise ValueError("means3d must be [N, 3]")
    if scales.shape != means3d.shape or quats.shape != (means3d.shape[0], 4):
        raise ValueError("scales must be [N,3] and quats [N,4]")
    return _ProjectGaussians.apply(means3d, scales, glob_scale, quats, viewmat, projmat, fx, fy,
                                   cx, cy, img_height, img_width, tuple(tile_bounds), clip_thresh)
