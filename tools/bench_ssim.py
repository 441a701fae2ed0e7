"""Fused SSIM (fwd+bwd) vs a torch conv2d implementation on a 1080p RGB image."""
import json, os, sys, torch, torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from tinysplat_b200.ssim import SSIM
H, W = 1080, 1920
img = torch.rand(H, W, 3, device="cuda", requires_grad=True)
gt = torch.rand(H, W, 3, device="cuda")
coords = torch.arange(11, dtype=torch.float32, device="cuda") - 5
g = torch.exp(-(coords ** 2) / (2 * 1.5 ** 2)); g = g / g.sum()
def filt(x):
    x = F.conv2d(x, g.view(1, 1, -1, 1).repeat(3, 1, 1, 1), groups=3)
    return F.conv2d(x, g.view(1, 1, 1, -1).repeat(3, 1, 1, 1), groups=3)
def torch_ssim(X, Y, C1=1e-4, C2=9e-4):
    m1, m2 = filt(X), filt(Y)
    s1, s2, s12 = filt(X * X) - m1 * m1, filt(Y * Y) - m2 * m2, filt(X * Y) - m1 * m2
    return (((2 * m1 * m2 + C1) / (m1 * m1 + m2 * m2 + C1)) * ((2 * s12 + C2) / (s1 + s2 + C2))).mean()
fused = SSIM(data_range=1.0, channel=3)
def run(fn, iters=20):
    for _ in range(3):
        img.grad = None; (1 - fn(img.permute(2, 0, 1)[None], gt.permute(2, 0, 1)[None])).backward()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        img.grad = None; (1 - fn(img.permute(2, 0, 1)[None], gt.permute(2, 0, 1)[None])).backward()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
a, b = run(fused), run(torch_ssim)
v1 = float(fused(img.permute(2, 0, 1)[None], gt.permute(2, 0, 1)[None])); v2 = float(torch_ssim(img.permute(2, 0, 1)[None], gt.permute(2, 0, 1)[None]))
print(json.dumps({"image": [H, W, 3], "fused_ssim_fwd_bwd_ms": round(a, 4), "torch_conv2d_ssim_fwd_bwd_ms": round(b, 4),
                  "speedup": round(b / a, 1), "value_fused": v1, "value_torch": v2}))
