"""Runs the UNMODIFIED training loop of the reference — `train()` of /root/reference/scripts/train.py with
the reference's own GaussianModel, Scene, Camera and GaussianRasterizer — for a few steps on CPU, with
the oracle standing in for the absent `gsplat` package and small stand-ins for the third-party modules
this image lacks (torchmetrics, pytorch_msssim, pytorch3d, plyfile, pycolmap, viser, websockets).

What it pins: "scripts/train.py runs unchanged on top of the five symbols" is EXERCISED, not asserted —
the argument order and return arities of the three calls, `xys.retain_grad()` + `extras['xys'].grad`
read by `update_grad_accum`, `backward(retain_graph=True)`, the optimizer over the six named parameter
groups, SH degree bookkeeping, and (second case) `densify_and_prune` really cloning / splitting / pruning
with its optimizer-state surgery, so that N changes between two renders.  The stand-in is the oracle
(same signatures and semantics as the `gsplat/` package of this repo, which needs a GPU); nothing here
touches the CUDA library.  Skipped where the reference checkout does not exist (the GPU box).
TEST INFRASTRUCTURE ONLY."""
import asyncio
import importlib
import importlib.util
import os
import sys
import types

import numpy as np
import pytest
import torch

import oracle
from oracle import ssim_oracle

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "scripts", "train.py")),
                                reason="reference checkout not present")

_STUBBED = ("gsplat", "tinysplat", "torchmetrics", "pytorch_msssim", "pytorch3d", "plyfile", "pycolmap", "viser",
            "websockets", "reference_train")


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _PSNR:
    def __init__(self, data_range=1.0):
        self.data_range = data_range

    def __call__(self, a, b):
        mse = ((a.detach() - b.detach()) ** 2).mean().clamp_min(1e-12)
        return 10.0 * torch.log10(self.data_range ** 2 / mse)


class _SSIM(torch.nn.Module):
    """pytorch_msssim.SSIM(data_range=1.0, size_average=True, channel=3) over the repo's CPU SSIM oracle."""
    def __init__(self, data_range=1.0, size_average=True, channel=3):
        super().__init__()
        self.data_range = data_range

    def forward(self, X, Y):
        return ssim_oracle.ssim(X, Y, data_range=self.data_range)


def _knn_points(p1, p2, K=1, **_):
    d = torch.cdist(p1, p2) ** 2
    dists, idx = d.topk(K, dim=-1, largest=False)
    return types.SimpleNamespace(dists=dists, idx=idx, knn=None)


def _ball_query(*a, **k):
    raise NotImplementedError("ball_query is imported by the reference but never called while training")


@pytest.fixture()
def reference():
    saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] in _STUBBED}
    for k in saved:
        sys.modules.pop(k, None)
    try:
        gsh = _module("gsplat.sh", spherical_harmonics=oracle.spherical_harmonics, num_sh_bases=oracle.num_sh_bases,
                      deg_from_sh=oracle.deg_from_sh)
        _module("gsplat", sh=gsh, project_gaussians=oracle.project_gaussians,
                rasterize_gaussians=oracle.rasterize_gaussians)
        tmi = _module("torchmetrics.image", PeakSignalNoiseRatio=_PSNR, StructuralSimilarityIndexMeasure=_PSNR)
        _module("torchmetrics", image=tmi)
        _module("pytorch_msssim", SSIM=_SSIM)
        ops = _module("pytorch3d.ops", knn_points=_knn_points, ball_query=_ball_query)
        _module("pytorch3d", ops=ops)
        _module("plyfile", PlyData=object, PlyElement=object)
        _module("pycolmap")
        vtf = _module("viser.transforms")
        _module("viser", transforms=vtf)
        _module("websockets")
        sys.path.insert(0, REF)
        tinysplat = importlib.import_module("tinysplat")          # the real package __init__
        spec = importlib.util.spec_from_file_location("reference_train", os.path.join(REF, "scripts", "train.py"))
        train_mod = importlib.util.module_from_spec(spec)
        sys.modules["reference_train"] = train_mod
        spec.loader.exec_module(train_mod)
        yield tinysplat, train_mod
    finally:
        if REF in sys.path:
            sys.path.remove(REF)
        for k in [k for k in sys.modules if k.split(".")[0] in _STUBBED]:
            sys.modules.pop(k, None)
        sys.modules.update(saved)


@pytest.mark.parametrize("densify", [False, True])
def test_unmodified_reference_training_loop_runs_on_the_five_symbols(reference, densify):
    tinysplat, train_mod = reference
    from tinysplat.scene import Camera, PointCloud, Scene
    torch.manual_seed(0)
    np.random.seed(0)
    dev = torch.device("cpu")
    W, H, n = 64, 48, 150
    argv = ["--train", "--device", "cpu", "--max-iter", "4", "--sh-degree", "2", "--sh-increment-interval", "2",
            "--warmup-grad", "2"]
    if densify:   # clone / split / prune + the optimizer-state surgery at step 2: N changes between steps
        argv += ["--warmup-densify", "2", "--interval-densify", "2", "--tau-means", "0.0"]
    args = train_mod.arg_parser().parse_args(argv)
    args.viewer = False
    # a blob of points in front of two cameras that look down +z
    xyz = torch.randn(n, 3) * 0.35 + torch.tensor([0.0, 0.0, 3.0])
    pcd = PointCloud(torch.arange(n), xyz, torch.randint(0, 256, (n, 3)).float(), torch.zeros(n))
    model = tinysplat.GaussianModel.from_pcd(pcd, **vars(args)).to(dev)
    model.device = dev
    fov = 0.9
    cams = []
    for k in range(2):
        img = torch.randint(0, 256, (H, W, 3)).float()            # LazyTensorImage divides a Tensor by 255
        cams.append(Camera(position=np.array([0.15 * k, 0.0, 0.0]), f_x=W / (2 * np.tan(fov / 2)),
                           f_y=H / (2 * np.tan(fov / 2)), fov_x=fov, fov_y=fov,
                           quat=np.array([1.0, 0.0, 0.0, 0.0]), near=0.01, far=100.0, image=img, name=f"cam{k}",
                           device=dev))
    rast = tinysplat.GaussianRasterizer(model, cams, device=dev)
    scene = Scene(cams, model, rast)
    before = {k: getattr(model, k).detach().clone() for k in ("means", "scales", "quats", "opacities", "colors_dc")}

    calls = {"render": 0, "grad_accum": 0}
    render = scene.render
    accum = model.update_grad_accum

    def counted_render(camera, dims=None):
        calls["render"] += 1
        img, extras = render(camera, dims)
        assert img.shape == (H, W, 3) and extras["depth"].shape == (H, W) and extras["xys"].requires_grad
        return img, extras

    def counted_accum(step, extras):
        calls["grad_accum"] += 1
        assert extras["xys"].grad is not None and extras["xys"].grad.shape[1] == 2
        return accum(step, extras)

    scene.render, model.update_grad_accum = counted_render, counted_accum
    asyncio.run(train_mod.train(model, scene, args))

    assert calls == {"render": 4, "grad_accum": 4}
    assert model.active_sh_degree == 2                            # 1 -> 2 at step 2, capped at --sh-degree
    for k, old in before.items():
        assert torch.isfinite(getattr(model, k)).all(), k
    if densify:
        assert model.means.shape[0] != n                          # the model was rebuilt with a different N ...
        assert model.colors_rest.shape[0] == model.means.shape[0] == model.means_grad_accum.shape[0]
    else:
        assert float(model.means_grad_accum.abs().sum()) > 0      # fed from extras['xys'].grad from step 2 on
        for k, old in before.items():                             # Adam moved every parameter group
            assert not torch.equal(getattr(model, k).detach(), old), k
