"""GPU-side breakdown of one bench step with torch.profiler (glue vs our kernels)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from torch.profiler import profile, ProfilerActivity
from tinysplat_b200 import synthetic
from tinysplat_b200.rasterizer import GaussianRasterizer, ParamModel
import bench

wl = sys.argv[1] if len(sys.argv) > 1 else "synthetic_1M_1080p"
pipe = sys.argv[2] if len(sys.argv) > 2 else "fused"
N, W, H, deg, dw, fwd_only = bench.WORKLOADS[wl]
dev = torch.device("cuda:0")
model = ParamModel(synthetic.make_scene(N, W, H, seed=0), dev, deg)
rast = GaussianRasterizer(model, None, dev, pipe)
gt = torch.rand(H, W, 3, device=dev)

def step(i):
    cam = bench.view_for(i, 0, W, H)
    img, ex = rast(cam, (W, H), deg)
    loss = (img - gt).abs().mean()
    loss.backward()
    model.zero_grad()

for i in range(5):
    step(i)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for i in range(5, 10):
        step(i)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=60))
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=30, max_name_column_width=60))
