"""Runs the UNMODIFIED reference adapter (/root/reference/tinysplat/splatting/rasterize.py) on
CPU with the oracle standing in for the absent gsplat package, and checks that the oracle's own
restatement of the adapter (oracle.render_reference_adapter, which tests, smoke and the CPU
baseline use) computes the same thing.  This anchors argument order, H-before-W, tuple arities
and the camera conventions on the reference's real call sites.  Skipped where the reference
checkout does not exist (the GPU box)."""
import importlib
import os
import sys
import types

import pytest
import torch

import oracle
from tinysplat_b200 import synthetic

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "tinysplat")),
                                reason="reference checkout not present")


@pytest.fixture()
def reference_rasterize():
    saved = {k: sys.modules.get(k) for k in list(sys.modules)
             if k == "gsplat" or k.startswith("gsplat.") or k == "tinysplat" or k.startswith("tinysplat.")}
    for k in saved:
        sys.modules.pop(k, None)
    # oracle-backed stand-in for the five symbols
    g = types.ModuleType("gsplat")
    gsh = types.ModuleType("gsplat.sh")
    gsh.spherical_harmonics = oracle.spherical_harmonics
    gsh.num_sh_bases = oracle.num_sh_bases
    gsh.deg_from_sh = oracle.deg_from_sh
    g.sh = gsh
    g.project_gaussians = oracle.project_gaussians
    g.rasterize_gaussians = oracle.rasterize_gaussians
    sys.modules["gsplat"], sys.modules["gsplat.sh"] = g, gsh
    # the reference package, without executing its __init__ (which needs pycolmap etc.)
    pkg = types.ModuleType("tinysplat")
    pkg.__path__ = [os.path.join(REF, "tinysplat")]
    sys.modules["tinysplat"] = pkg
    sub = types.ModuleType("tinysplat.splatting")
    sub.__path__ = [os.path.join(REF, "tinysplat", "splatting")]
    sys.modules["tinysplat.splatting"] = sub
    mg = types.ModuleType("tinysplat.splatting.model_gaussian")   # real one needs pytorch3d, plyfile ...
    mg.GaussianModel = type("GaussianModel", (), {})
    sys.modules["tinysplat.splatting.model_gaussian"] = mg
    try:
        mod = importlib.import_module("tinysplat.splatting.rasterize")
        yield mod
    finally:
        for k in [k for k in sys.modules if k == "gsplat" or k.startswith("gsplat.") or k == "tinysplat"
                  or k.startswith("tinysplat.")]:
            sys.modules.pop(k, None)
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v


class _Model:
    pass


@pytest.mark.parametrize("W,H,deg", [(128, 128, 3), (72, 40, 1)])
def test_unmodified_reference_adapter_equals_oracle_adapter(reference_rasterize, W, H, deg):
    cam = synthetic.make_camera(W, H, yaw_deg=4.0, shift=(0.1, 0.0, 0.0))
    sc = synthetic.make_scene(200, W, H, seed=3)
    sc["background"] = torch.tensor([0.1, 0.4, 0.9])
    names = ["means", "scales", "quats", "opacities", "colors_dc", "colors_rest"]
    m = _Model()
    for k in names:
        setattr(m, k, sc[k].clone().requires_grad_(True))
    m.background = sc["background"]
    m.active_sh_degree = deg
    rast = reference_rasterize.GaussianRasterizer(m, [cam], device=torch.device("cpu"))
    img, extras = rast(cam, None, deg)
    assert img.shape == (H, W, 3) and extras["depth"].shape == (H, W)
    (img.sum() + extras["depth"].sum()).backward()
    assert extras["xys"].grad is not None                      # what update_grad_accum reads
    p = {k: sc[k].clone().requires_grad_(True) for k in names}
    p["background"] = sc["background"]
    rimg, rex = oracle.render_reference_adapter(p, cam.view_matrix, cam.proj_matrix, cam.f_x, cam.f_y, (W, H), deg)
    (rimg.sum() + rex["depth"].sum()).backward()
    assert torch.equal(img, rimg) and torch.equal(extras["depth"], rex["depth"])
    assert torch.equal(extras["radii"], rex["radii"])
    assert torch.allclose(extras["xys"].grad, rex["xys"].grad, rtol=0, atol=0)
    for k in names:   # the reference applies sigmoid twice (two graph nodes): last-bit differences only
        assert torch.allclose(getattr(m, k).grad, p[k].grad, rtol=1e-5, atol=1e-6), k
