"""Host-side mirror of tinysplat's raster adapter, for callers that do not have the
reference checkout on their path (bench.py, smoke, the GPU tests).

`GaussianRasterizer(model, cameras, device)(camera, dims, sh_degree) -> (rgb[H,W,3], extras)`
has the constructor, call signature, outputs and `extras` keys of
tinysplat.splatting.rasterize.GaussianRasterizer  [REF tinysplat/splatting/rasterize.py:13-62].
`model` is anything with the GaussianModel parameter attributes
[REF tinysplat/splatting/model_gaussian.py:84-89] plus `background` and `active_sh_degree`.

Two pipelines over the same kernels:
  * "reference": the reference's exact op sequence through the five gsplat symbols —
    project, SH on concatenated coefficients, rasterise RGB, rasterise depth-as-colour.
  * "fused" (default): ONE autograd node (tinysplat_b200.fused.render_fused): the adapter's
    torch-side activations are folded into the kernels, RGB + depth share binning, sort and one
    4-channel blend pass, SH reads the (dc, rest) pair without the per-step
    concatenation.  Same numbers; ~11 kernel launches per forward+backward instead of ~70.
  * "unfused4": the previous composition of the five public ops with a 4-channel rasterise
    (kept as a parity cross-check of the fused node).
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import torch
from torch import Tensor

from .fused import render_fused
from .project import project_gaussians
from .rasterize import rasterize_gaussians
from .sh import spherical_harmonics, spherical_harmonics_split

TILE = 16


def tile_grid(width: int, height: int) -> Tuple[int, int, int]:
    """(ceil(W/16), ceil(H/16), 1)  [REF rasterize.py:88-94]."""
    return (-(-width // TILE), -(-height // TILE), 1)


class GaussianRasterizer:
    def __init__(self, model, cameras: Optional[Sequence] = None, device="cuda:0",
                 pipeline: str = "fused"):
        if pipeline not in ("fused", "reference", "unfused4"):
            raise ValueError("pipeline must be 'fused', 'reference' or 'unfused4'")
        self.model = model
        self.device = torch.device(device)
        self.pipeline = pipeline
        # data-parallel training: a tinysplat_b200.parallel.PackedGradExchange makes the fused
        # node's backward exchange packed gradient rows and return already-reduced gradients
        self.grad_exchange = None

    # -- shared front end: projection + view-dependent colour ---------------------------------
    def _project(self, camera, width: int, height: int):
        m = self.model
        view = camera.view_matrix.to(self.device)
        full_proj = camera.proj_matrix.to(self.device) @ view
        unit_q = m.quats / m.quats.norm(dim=-1, keepdim=True)
        return project_gaussians(m.means, torch.exp(m.scales), 1.0, unit_q, view[:3, :], full_proj,
                                 camera.f_x, camera.f_y, width / 2, height / 2, height, width,
                                 tile_grid(width, height)), view

    def _view_dirs(self, view: Tensor) -> Tensor:
        # the reference uses the view matrix's translation column as the eye position
        # [REF rasterize.py:77-79]; reproduced as is (SURVEY.md 8a note on a3)
        d = self.model.means - view[:3, 3]
        return d / d.norm(dim=-1, keepdim=True)

    def _render_fused(self, camera, width: int, height: int, sh_degree: int):
        m = self.model
        # one small H2D for both matrices (full projection composed on the host)
        view = camera.view_matrix.float()
        full = camera.proj_matrix.float() @ view
        exchange = self.grad_exchange if torch.is_grad_enabled() else None
        if exchange is None:
            mats = torch.stack([view, full]).to(self.device, non_blocking=True)
            view_d, full_d, cam_row = mats[0], mats[1], None
        else:
            # + the 32-float camera row of the data-parallel exchange, in the same small H2D copy
            host = torch.cat([view.reshape(-1), full.reshape(-1), view[:3].reshape(-1), full.reshape(-1),
                              torch.tensor([float(camera.f_x), float(camera.f_y), 0.0, 0.0])])
            dev_buf = host.to(self.device, non_blocking=True)
            view_d, full_d, cam_row = dev_buf[:16].view(4, 4), dev_buf[16:32].view(4, 4), dev_buf[32:64]
        rgb, depth_img, _, xys, _, radii = render_fused(
            m.means, m.scales, m.quats, m.opacities, m.colors_dc, m.colors_rest, view_d, full_d,
            camera.f_x, camera.f_y, width, height, sh_degree, m.background,
            grad_exchange=exchange, cam_row=cam_row)
        extras: Dict = {"depth": depth_img, "radii": radii, "xys": xys,
                        "camera": {"height": camera.height, "width": camera.width}}
        return rgb, extras      # already clamped to <= 1 inside the blend kernel

    def __call__(self, camera, dims: Optional[Tuple[int, int]], sh_degree: int):
        m = self.model
        width, height = dims if dims is not None else (camera.width, camera.height)
        if self.pipeline == "fused":
            return self._render_fused(camera, width, height, sh_degree)
        (xys, depths, radii, conics, num_tiles, _), view = self._project(camera, width, height)
        if xys.requires_grad:
            xys.retain_grad()
        dirs = self._view_dirs(view)
        opac = torch.sigmoid(m.opacities)
        bg = m.background.to(self.device)

        if self.pipeline == "reference":
            coeffs = torch.cat([m.colors_dc[:, None, :], m.colors_rest], dim=1)
            rgbs = torch.clamp(spherical_harmonics(sh_degree, dirs, coeffs) + 0.5, min=0.0)
            img, _ = rasterize_gaussians(xys, depths, radii, conics, num_tiles, rgbs, opac,
                                         height, width, bg)
            img = torch.clamp(img, max=1.0)
            dimg, _ = rasterize_gaussians(xys, depths, radii, conics, num_tiles,
                                          depths[:, None].repeat(1, 3), opac, height, width, bg)
            depth_img = dimg[:, :, 0]
        else:
            rgbs = torch.clamp(spherical_harmonics_split(sh_degree, dirs, m.colors_dc,
                                                         m.colors_rest) + 0.5, min=0.0)
            rgbd = torch.cat([rgbs, depths[:, None]], dim=1)
            out, _ = rasterize_gaussians(xys, depths, radii, conics, num_tiles, rgbd, opac,
                                         height, width, torch.cat([bg, bg[:1]]))
            img = torch.clamp(out[:, :, :3], max=1.0)
            depth_img = out[:, :, 3]

        extras: Dict = {"depth": depth_img, "radii": radii, "xys": xys,
                        "camera": {"height": camera.height, "width": camera.width}}
        return img, extras


class ParamModel:
    """Minimal stand-in for GaussianModel: holds the six Parameters + background + degree."""

    def __init__(self, params: Dict[str, Tensor], device="cuda:0", sh_degree: int = 3,
                 requires_grad: bool = True):
        dev = torch.device(device)
        for name in ("means", "scales", "quats", "opacities", "colors_dc", "colors_rest"):
            t = params[name].detach().to(dev).float().contiguous()
            setattr(self, name, t.requires_grad_(requires_grad))
        self.background = params["background"].detach().to(dev).float()
        self.active_sh_degree = sh_degree
        self.device = dev

    def parameters(self):
        return [self.means, self.scales, self.quats, self.opacities, self.colors_dc, self.colors_rest]

    def zero_grad(self):
        for p in self.parameters():
            p.grad = None
