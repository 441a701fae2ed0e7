"""Fused Adam (ts_adam_step) vs torch.optim.Adam on the 1M-Gaussian parameter set: ms/step and
achieved HBM GB/s (28 B per element: param/grad/exp_avg/exp_avg_sq read, param + 2 moments written)."""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from tinysplat_b200.optim import FusedAdam
from tinysplat_b200 import synthetic
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
sc = synthetic.make_scene(N, 1920, 1080, seed=0)
names = ["means", "colors_dc", "colors_rest", "scales", "quats", "opacities"]
lrs = dict(means=0.00016, colors_dc=0.0025, colors_rest=0.000125, scales=0.005, quats=0.001, opacities=0.05)
def make():
    ps = {k: torch.nn.Parameter(sc[k].clone().cuda()) for k in names}
    for p in ps.values():
        p.grad = torch.randn_like(p) * 1e-3
    return ps, [{"params": [ps[k]], "lr": lrs[k], "name": k} for k in names]
def bench(opt, iters=30):
    for _ in range(5): opt.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): opt.step()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
elems = sum(sc[k].numel() for k in names)
out = {"gaussians": N, "elements": elems, "algorithmic_bytes": 28 * elems}
for label, ctor in [("ts_adam_step (1 launch)", lambda g: FusedAdam(g)),
                    ("torch Adam foreach", lambda g: torch.optim.Adam(g, foreach=True)),
                    ("torch Adam fused", lambda g: torch.optim.Adam(g, fused=True)),
                    ("torch Adam single-tensor", lambda g: torch.optim.Adam(g, foreach=False))]:
    ps, groups = make()
    ms = bench(ctor(groups))
    out[label] = {"ms": round(ms, 4), "GB/s": round(28 * elems / ms / 1e6, 1)}
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
out["hbm_peak_GB/s"] = peak
out["frac_of_peak"] = round(out["ts_adam_step (1 launch)"]["GB/s"] / peak, 3)
print(json.dumps(out))
