"""Drop-in `gsplat` package: the exact five-symbol surface tinysplat imports
[REF tinysplat/splatting/rasterize.py:3-4; tinysplat/splatting/model_gaussian.py:14],
backed by tinysplat_b200's sm_100a kernels.  Put the repo root on sys.path (ahead of any other
gsplat) and tinysplat's scripts/train.py runs unchanged; see INTEGRATION.md."""
from tinysplat_b200.project import project_gaussians      # noqa: F401
from tinysplat_b200.rasterize import rasterize_gaussians  # noqa: F401
from . import sh                                            # noqa: F401

__version__ = "0.1.3+tinysplat_b200"
