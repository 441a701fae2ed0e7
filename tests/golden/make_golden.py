"""Generates tests/golden/config1.npz: BASELINE config 1 (256 synthetic Gaussians, 128x128)
rendered by the fp64 oracle through the reference adapter's op sequence, with gradients of a
fixed scalar loss.  Run from the repo root:  python tests/golden/make_golden.py

PARITY UNPINNED upstream (no reference fixtures exist): these vectors pin the CUDA path to
THIS repo's stated oracle, and pin the oracle against accidental change."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from tinysplat_b200 import synthetic  # noqa: E402

W = H = 128
N = 256
PARAMS = ["means", "scales", "quats", "opacities", "colors_dc", "colors_rest"]


def loss_weights(H, W):
    g = torch.Generator().manual_seed(1234)
    return torch.rand(H, W, 3, generator=g, dtype=torch.float64), torch.rand(H, W, generator=g, dtype=torch.float64)


def run(dtype):
    cam = synthetic.make_camera(W, H)
    sc = synthetic.make_scene(N, W, H, seed=0)
    sc["background"] = torch.tensor([0.2, 0.5, 0.8])
    p = {k: v.to(dtype) for k, v in sc.items()}
    for k in PARAMS:
        p[k].requires_grad_(True)
    img, ex = oracle.render_reference_adapter(p, cam.view_matrix, cam.proj_matrix, cam.f_x, cam.f_y,
                                              (W, H), 3)
    wi, wd = loss_weights(H, W)
    loss = (img * wi.to(dtype)).sum() + 0.1 * (ex["depth"] * wd.to(dtype)).sum()
    loss.backward()
    out = {"img": img, "depth": ex["depth"], "xys": ex["xys"], "radii": ex["radii"],
           "v_xys": ex["xys"].grad}
    for k in PARAMS:
        out["v_" + k] = p[k].grad
    return sc, {k: v.detach() for k, v in out.items()}


if __name__ == "__main__":
    sc, o64 = run(torch.float64)
    _, o32 = run(torch.float32)
    print("fp32-oracle vs fp64-oracle (calibration for the CUDA tolerances):")
    for k in o64:
        a, b = o64[k].double(), o32[k].double()
        err = (a - b).abs().max().item()
        ref = a.abs().max().item()
        print(f"  {k:14s} max|d|={err:.3e}  max|ref|={ref:.3e}  rel={err / max(ref, 1e-30):.3e}")
    save = {"in_" + k: v.numpy() for k, v in sc.items()}
    save.update({k: v.to(torch.float32 if v.is_floating_point() else v.dtype).numpy() for k, v in o64.items()})
    np.savez_compressed(os.path.join(os.path.dirname(__file__), "config1.npz"), **save)
    print("wrote config1.npz")
