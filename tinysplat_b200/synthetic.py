"""Seeded synthetic scenes of the shape SURVEY.md section 8(d) names (no dataset on disk: the
T&T "truck" configs use this stand-in and are labelled synthetic everywhere).

Camera conventions follow the reference: world->camera `view`, +z forward, projection
matrix as built by Camera.update_proj_matrix  [REF tinysplat/scene.py:96-121]; near/far as
the dataset loader sets them  [REF tinysplat/dataset.py:91-92]."""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Tuple

import torch


@dataclass
class SynthCamera:
    """Duck-types the fields of tinysplat.scene.Camera that the raster adapter reads
    [REF tinysplat/splatting/rasterize.py:64-73]."""
    view_matrix: torch.Tensor   # [4,4] fp32 CPU, world -> camera
    proj_matrix: torch.Tensor   # [4,4] fp32 CPU
    f_x: float
    f_y: float
    width: int
    height: int


def make_camera(width: int, height: int, fov_x_deg: float = 60.0, yaw_deg: float = 0.0,
                shift: Tuple[float, float, float] = (0.0, 0.0, 0.0),
                znear: float = 0.001, zfar: float = 1000.0) -> SynthCamera:
    fov_x = math.radians(fov_x_deg)
    fx = 0.5 * width / math.tan(0.5 * fov_x)
    fy = fx
    fov_y = 2.0 * math.atan(0.5 * height / fy)
    a = math.radians(yaw_deg)
    R = torch.tensor([[math.cos(a), 0.0, -math.sin(a)],
                      [0.0, 1.0, 0.0],
                      [math.sin(a), 0.0, math.cos(a)]], dtype=torch.float64)
    p = torch.tensor(shift, dtype=torch.float64)
    V = torch.eye(4, dtype=torch.float64)
    V[:3, :3] = R
    V[:3, 3] = -R @ p
    P = torch.zeros(4, 4, dtype=torch.float64)
    P[0, 0] = 1.0 / math.tan(0.5 * fov_x)
    P[1, 1] = 1.0 / math.tan(0.5 * fov_y)
    P[2, 2] = (zfar + znear) / (zfar - znear)
    P[2, 3] = -zfar * znear / (zfar - znear)
    P[3, 2] = 1.0
    return SynthCamera(V.float(), P.float(), fx, fy, width, height)


def make_scene(num_points: int, width: int, height: int, sh_degree: int = 3, seed: int = 0,
               mean_radius_px: float = 6.0, fov_x_deg: float = 60.0,
               depth_range: Tuple[float, float] = (2.0, 10.0),
               dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """GaussianModel-shaped parameter dict  [REF tinysplat/splatting/model_gaussian.py:84-89]:
    means[N,3], scales[N,3] (log), quats[N,4] (w,x,y,z), opacities[N,1] (logit),
    colors_dc[N,3], colors_rest[N,K-1,3], background[3]."""
    g = torch.Generator().manual_seed(seed)
    N = num_points
    fov_x = math.radians(fov_x_deg)
    fx = 0.5 * width / math.tan(0.5 * fov_x)
    tan_x = math.tan(0.5 * fov_x)
    tan_y = 0.5 * height / fx
    z = torch.rand(N, generator=g, dtype=torch.float64) * (depth_range[1] - depth_range[0]) \
        + depth_range[0]
    u = torch.rand(N, 2, generator=g, dtype=torch.float64) * 2.1 - 1.05
    means = torch.stack([u[:, 0] * tan_x * z, u[:, 1] * tan_y * z, z], dim=-1)
    # radius = ceil(3*sqrt(sigma_px^2 + 0.3)); pick the world sigma whose projection at the
    # mean depth gives the requested mean radius (log-normal spread sigma=0.5)
    sigma_px = math.sqrt(max((mean_radius_px / 3.0) ** 2 - 0.3, 0.05))
    z_mid = 0.5 * (depth_range[0] + depth_range[1])
    mu = math.log(sigma_px * z_mid / fx) - 0.72   # empirical: E[1/z], max-of-3 lognormals, ceil
    scales = mu + 0.5 * torch.randn(N, 3, generator=g, dtype=torch.float64)
    uvw = torch.rand(N, 3, generator=g, dtype=torch.float64)
    quats = torch.stack([
        torch.sqrt(1 - uvw[:, 0]) * torch.sin(2 * math.pi * uvw[:, 1]),
        torch.sqrt(1 - uvw[:, 0]) * torch.cos(2 * math.pi * uvw[:, 1]),
        torch.sqrt(uvw[:, 0]) * torch.sin(2 * math.pi * uvw[:, 2]),
        torch.sqrt(uvw[:, 0]) * torch.cos(2 * math.pi * uvw[:, 2]),
    ], dim=-1)
    opacities = 1.5 * torch.randn(N, 1, generator=g, dtype=torch.float64)
    K = (sh_degree + 1) ** 2
    sh = 0.3 * torch.randn(N, K, 3, generator=g, dtype=torch.float64)
    return {
        "means": means.to(dtype),
        "scales": scales.to(dtype),
        "quats": quats.to(dtype),
        "opacities": opacities.to(dtype),
        "colors_dc": sh[:, 0, :].contiguous().to(dtype),
        "colors_rest": sh[:, 1:, :].contiguous().to(dtype),
        "background": torch.zeros(3, dtype=dtype),
    }
