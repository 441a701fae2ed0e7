"""Full training step à la scripts/train.py at 1M Gaussians / 1080p: render, 0.8 L1 + 0.2 (1-SSIM),
backward, Adam.  Compares this repo's fused pieces against torch stand-ins for loss/optimizer."""
import json, os, sys, torch, torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from tinysplat_b200 import synthetic
from tinysplat_b200.optim import FusedAdam
from tinysplat_b200.rasterizer import GaussianRasterizer, ParamModel
from tinysplat_b200.ssim import SSIM
import bench
N, W, H = 1_000_000, 1920, 1080
dev = torch.device("cuda:0")
names = ["means", "scales", "quats", "opacities", "colors_dc", "colors_rest"]
lrs = dict(means=0.00016, colors_dc=0.0025, colors_rest=0.000125, scales=0.005, quats=0.001, opacities=0.05)
gt = torch.rand(H, W, 3, device=dev)
coords = torch.arange(11, dtype=torch.float32, device=dev) - 5
gw = torch.exp(-(coords ** 2) / 4.5); gw = gw / gw.sum()
def filt(x):
    x = F.conv2d(x, gw.view(1, 1, -1, 1).repeat(3, 1, 1, 1), groups=3)
    return F.conv2d(x, gw.view(1, 1, 1, -1).repeat(3, 1, 1, 1), groups=3)
def torch_ssim(X, Y, C1=1e-4, C2=9e-4):
    m1, m2 = filt(X), filt(Y)
    s1, s2, s12 = filt(X * X) - m1 * m1, filt(Y * Y) - m2 * m2, filt(X * Y) - m1 * m2
    return (((2 * m1 * m2 + C1) / (m1 * m1 + m2 * m2 + C1)) * ((2 * s12 + C2) / (s1 + s2 + C2))).mean()
def run(use_fused_loss_opt, pipeline="fused", steps=20):
    model = ParamModel(synthetic.make_scene(N, W, H, seed=0), dev, 3)
    params = {k: torch.nn.Parameter(getattr(model, k).detach()) for k in names}
    for k, p in params.items(): setattr(model, k, p)
    groups = [{"params": [params[k]], "lr": lrs[k], "name": k} for k in names]
    opt = FusedAdam(groups) if use_fused_loss_opt else torch.optim.Adam(groups)
    ssim = SSIM(data_range=1.0, channel=3) if use_fused_loss_opt else torch_ssim
    rast = GaussianRasterizer(model, None, dev, pipeline)
    def step(i):
        img, ex = rast(bench.view_for(i, 0, W, H), None, 3)
        loss = 0.8 * (img - gt).abs().mean() + 0.2 * (1 - ssim(img.permute(2, 0, 1)[None], gt.permute(2, 0, 1)[None]))
        loss.backward(); opt.step(); opt.zero_grad(set_to_none=True)
    for i in range(5): step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps): step(5 + i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps
out = {"workload": "1M Gaussians, 1080p, full train step (render + L1 + DSSIM + backward + Adam)"}
out["fused adapter + fused SSIM + FusedAdam, ms/step"] = round(run(True), 3)
out["fused adapter + torch conv2d SSIM + torch Adam, ms/step"] = round(run(False), 3)
out["drop-in gsplat ops + torch conv2d SSIM + torch Adam, ms/step"] = round(run(False, "reference"), 3)
print(json.dumps(out))
