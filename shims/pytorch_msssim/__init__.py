"""Import-name stand-in for the `pytorch_msssim` package as tinysplat uses it
(`from pytorch_msssim import SSIM`  [REF tinysplat/splatting/model_gaussian.py:12,57]).

Opt-in: put `<repo>/shims` on `sys.path`.  `SSIM` is tinysplat_b200's fused sm_100a
implementation (tinysplat_b200/ssim.py); there is no CPU path."""
from tinysplat_b200.ssim import SSIM  # noqa: F401

__all__ = ["SSIM"]
