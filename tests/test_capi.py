"""CPU checks of the C-ABI boundary: the library builds, loads, exports exactly the symbols
include/tinysplat_b200.h declares, and rejects bad arguments without touching a GPU."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "tinysplat_b200.h")


def header_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"TS_API\s+[\w\s\*]+?\b(ts_\w+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    syms = header_symbols()
    for must in ("ts_project_fwd", "ts_project_bwd", "ts_sh_fwd", "ts_sh_bwd", "ts_bin_count",
                 "ts_bin_scan", "ts_bin_emit", "ts_bin_sort", "ts_blend_fwd", "ts_blend_bwd",
                 "ts_blend_unpack_grads"):
        assert must in syms


def test_library_exports_every_declared_symbol(lib):
    from tinysplat_b200 import _lib
    declared = header_symbols()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    # and the ctypes table binds exactly the header's set
    assert sorted(_lib.exported_symbols()) == declared


def test_no_torch_or_python_in_the_abi(lib):
    from tinysplat_b200 import _lib
    import subprocess
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "libtorch" not in out and "libpython" not in out and "libc10" not in out


def test_version_and_record_sizes(lib):
    assert lib.ts_version() >= 100
    assert lib.ts_rec_floats() == 12 and lib.ts_grad_floats() == 12
    assert lib.ts_bin_smem_sort_cap() >= 4096


def test_invalid_arguments_are_rejected_before_any_cuda_call(lib):
    N = None
    assert lib.ts_project_fwd(-1, N, N, 1.0, N, N, N, 1.0, 1.0, 0.0, 0.0, 16, 16, 1, 1, 0.01, 0,
                              N, N, N, N, N, N, N, 0, N, N, N) == -1
    assert lib.ts_sh_fwd(4, 5, 16, N, N, N, N, N, 3, N, N, 0, N) == -1        # degree > 4
    assert lib.ts_sh_fwd(4, 3, 9, N, N, N, N, N, 3, N, N, 0, N) == -1         # K too small
    assert lib.ts_blend_fwd(7, 16, 16, 1, 1, N, N, N, N, N, N, N, N, 0, 0, N, N) == -1  # 7 channels
    assert lib.ts_bin_scan(0, N, N, N, 1, N) == -1
    # N == 0 is a valid no-op everywhere
    assert lib.ts_project_bwd(0, N, N, 1.0, N, N, N, 1.0, 1.0, 0.0, 0.0, 16, 16, 0,
                              N, N, N, N, N, N, N, N, N, N, N, N) == 0
    assert lib.ts_sh_bwd(0, 3, 16, N, N, N, 3, N, N, N, 0, N) == 0


def test_misaligned_pointer_is_reported(lib):
    buf = (ctypes.c_float * 64)()
    base = ctypes.addressof(buf)
    base += (16 - base % 16) % 16
    bad = ctypes.c_void_p(base + 4)
    ok = ctypes.c_void_p(base)
    assert lib.ts_sh_fwd(1, 0, 1, bad, None, ok, None, ok, 3, None, None, 0, None) == -2


def test_product_fails_loudly_without_cuda():
    import gsplat
    from tinysplat_b200._lib import TinysplatError
    x = torch.zeros(4, 3)
    with pytest.raises(TinysplatError):
        gsplat.project_gaussians(x, x, 1.0, torch.zeros(4, 4), torch.eye(4)[:3], torch.eye(4), 1., 1., 8., 8.,
                                 16, 16, (1, 1, 1))
    with pytest.raises(TinysplatError):
        gsplat.sh.spherical_harmonics(0, x, torch.zeros(4, 1, 3))
    with pytest.raises(TinysplatError):
        gsplat.rasterize_gaussians(torch.zeros(4, 2), torch.zeros(4), torch.zeros(4, dtype=torch.int32), x,
                                   torch.zeros(4, dtype=torch.int32), x, torch.zeros(4, 1), 16, 16, torch.zeros(3))


def test_fused_adam_fails_loudly_on_cpu_parameters():
    from tinysplat_b200.optim import FusedAdam
    from tinysplat_b200._lib import TinysplatError
    p = torch.nn.Parameter(torch.zeros(4, 3))
    opt = FusedAdam([{"params": [p], "lr": 0.1, "name": "means"}])
    p.grad = torch.ones_like(p)
    with pytest.raises(TinysplatError):
        opt.step()


def test_product_never_imports_the_oracle():
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r); import gsplat, tinysplat_b200, tinysplat_b200.rasterizer, "
            "tinysplat_b200.parallel, tinysplat_b200.optim, tinysplat_b200.ssim, tinysplat_b200.knn; assert not any(m.split('.')[0] == 'oracle' for m in sys.modules), 'oracle imported'"
            % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_gsplat_surface():
    import gsplat
    from gsplat.sh import spherical_harmonics, num_sh_bases, deg_from_sh
    from gsplat import project_gaussians, rasterize_gaussians
    assert [num_sh_bases(d) for d in range(5)] == [1, 4, 9, 16, 25]
    assert [deg_from_sh(n) for n in (1, 4, 9, 16, 25)] == [0, 1, 2, 3, 4]
    with pytest.raises(ValueError):
        deg_from_sh(7)


def _rowmask_case(lib, x, y, cov, op, tile_x, tile_y):
    """Brute force: every pixel of the tile the blend loop would accept (fp32 arithmetic in the
    kernels' operation order, without fma — the difference is inside the mask's noise margin)
    must have its (row, half) bit set in ts_debug_rowmask."""
    import numpy as np
    f = np.float32
    a, b, c = cov
    det = a * c - b * b
    con = (c / det, -b / det, a / det)
    LOG2E = 1.4426950408889634
    q1 = np.array([0.5 * LOG2E * con[0], LOG2E * con[1], 0.5 * LOG2E * con[2], op], dtype=f)
    tau = np.log(255.0 * op) + 0.01
    hx = np.sqrt(2 * tau * a) * 1.001 + 0.01     # cov_xx = a  (cov = conic^-1)
    hy = np.sqrt(2 * tau * c) * 1.001 + 0.01
    q0 = np.array([x, y, hx, hy], dtype=f)
    mask = lib.ts_debug_rowmask(q0.ctypes.data_as(ctypes.c_void_p), q1.ctypes.data_as(ctypes.c_void_p),
                                tile_x, tile_y)
    px = (f(tile_x * 16) + np.arange(16, dtype=f) + f(0.5))[None, :]
    py = (f(tile_y * 16) + np.arange(16, dtype=f) + f(0.5))[:, None]
    dx = q0[0] - px
    dy = q0[1] - py
    t = q1[1] * dy + q1[0] * dx
    pw = t * dx + (q1[2] * dy) * dy
    with np.errstate(over="ignore", invalid="ignore"):
        alpha = np.minimum(f(0.999), q1[3] * np.exp2(-pw.astype(np.float64)).astype(f))
    lit = (pw >= 0) & (alpha >= f(1.0 / 255.0))
    bits = np.zeros((16, 2), dtype=bool)
    for row in range(16):
        for half in range(2):
            bits[row, half] = (mask >> (2 * row + half)) & 1
    need = lit.reshape(16, 2, 8).any(axis=2)
    return need, bits


def test_exact_rowmask_is_a_superset_of_lit_pixels_and_tight(lib):
    import numpy as np
    rng = np.random.default_rng(0)
    missed = 0
    n_need = n_bits = 0
    for it in range(3000):
        kind = it % 3
        if kind == 0:      # bench-like blobs
            s1, s2 = np.exp(rng.normal(0.6, 0.5, 2))
        elif kind == 1:    # needles: one axis huge, one at the blur floor
            s1, s2 = np.exp(rng.uniform(3, 7)), np.exp(rng.uniform(-3, 0))
        else:              # big soft splats
            s1, s2 = np.exp(rng.uniform(2, 5, 2))
        th = rng.uniform(0, np.pi)
        R = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
        S = R @ np.diag([s1 * s1, s2 * s2]) @ R.T
        cov = (S[0, 0] + 0.3, S[0, 1], S[1, 1] + 0.3)
        op = float(np.clip(1 / (1 + np.exp(-rng.normal(0, 1.5))), 1.0 / 255 + 1e-4, 0.9999))
        tile_x, tile_y = int(rng.integers(0, 120)), int(rng.integers(0, 68))
        reach = 3.5 * max(s1, s2) if kind != 1 else rng.uniform(0, 3.0 * max(s1, s2))
        x = tile_x * 16 + 8 + rng.uniform(-1, 1) * (8 + reach)
        y = tile_y * 16 + 8 + rng.uniform(-1, 1) * (8 + reach)
        need, bits = _rowmask_case(lib, x, y, cov, op, tile_x, tile_y)
        missed += int((need & ~bits).sum())
        if kind == 0:
            n_need += int(need.sum())
            n_bits += int(bits.sum())
    assert missed == 0, f"{missed} lit (row, half) cells are not covered by the exact row mask"
    # tightness on bench-like Gaussians: the mask is exact up to its safety margins
    assert n_need > 1000 and n_bits <= 1.10 * n_need, (n_need, n_bits)


def test_rowmask_sentinels(lib):
    import numpy as np
    f = np.float32
    # a conic that is not positive along x (broken covariance) or NaN: no culling at all
    q0 = np.array([100.0, 40.0, 5.0, 5.0], dtype=f)
    for A in (0.0, -0.3, float("nan")):
        q1 = np.array([A, 0.1, 0.2, 0.5], dtype=f)
        assert lib.ts_debug_rowmask(q0.ctypes.data_as(ctypes.c_void_p), q1.ctypes.data_as(ctypes.c_void_p),
                                    0, 0) == 0xffffffff
    q1 = np.array([0.1, 0.0, 0.1, 0.5], dtype=f)
    for hx, expect in ((1e30, 0xffffffff), (-1e30, 0)):
        q0 = np.array([8.0, 8.0, hx, hx], dtype=f)
        assert lib.ts_debug_rowmask(q0.ctypes.data_as(ctypes.c_void_p), q1.ctypes.data_as(ctypes.c_void_p),
                                    0, 0) == expect


def test_optional_import_name_shims_resolve_to_the_fused_ops():
    """shims/ provides the import names of two of the reference's absent third-party dependencies
    (pytorch_msssim.SSIM, pytorch3d.ops.knn_points) on top of the fused kernels; opt-in via sys.path."""
    import subprocess
    import sys
    code = ("import sys; sys.path[:0] = [%r, %r]; "
            "from pytorch_msssim import SSIM; from pytorch3d.ops import knn_points, ball_query; "
            "import tinysplat_b200.ssim as s, tinysplat_b200.knn as k; "
            "assert SSIM is s.SSIM and knn_points is k.knn_points; "
            "m = SSIM(data_range=1.0, size_average=True, channel=3)" % (ROOT, os.path.join(ROOT, "shims")))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
