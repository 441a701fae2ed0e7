"""CPU oracle for the tinysplat hot path (TEST INFRASTRUCTURE ONLY).

PARITY UNPINNED: the reference (maxgillett/tinysplat) delegates this path's arithmetic to
the third-party ``gsplat`` package (legacy 0.1.x functional API, version not pinned by the
reference, not vendored, not installable here).  The reference itself ships no tests, golden
vectors or fixtures for the path.  This package restates the published algorithm (3DGS,
arXiv 2308.04079, + the gsplat-legacy constants listed in ``gsplat_oracle.py``) and is
pinned only by fp64 gradcheck, closed-form cases and invariances (tests/test_oracle.py).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` leg may import this package.  The product (``tinysplat_b200`` / ``gsplat``)
never does.
"""
from .gsplat_oracle import (  # noqa: F401
    project_gaussians,
    rasterize_gaussians,
    spherical_harmonics,
    num_sh_bases,
    deg_from_sh,
    render_reference_adapter,
)
