"""tinysplat_b200 — B200 (sm_100a) differentiable Gaussian rasterizer behind the five gsplat
symbols tinysplat imports.  Host side is Python/PyTorch plumbing over a C-ABI CUDA library
(include/tinysplat_b200.h); there is no CPU fallback."""
from .project import project_gaussians            # noqa: F401
from .rasterize import rasterize_gaussians        # noqa: F401
from .sh import spherical_harmonics, spherical_harmonics_split, num_sh_bases, deg_from_sh  # noqa: F401

__version__ = "0.1.0"
