"""gsplat.sh surface  [REF tinysplat/splatting/rasterize.py:3; model_gaussian.py:14]."""
from tinysplat_b200.sh import spherical_harmonics, num_sh_bases, deg_from_sh  # noqa: F401
