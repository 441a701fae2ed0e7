"""Multi-GPU check of the gradient-exchange strategies (run under torchrun, one rank per GPU):
renders one view per rank through the fused adapter with strategy "allreduce", "packed" (NCCL
all-to-all + shard backward + all-gather) and "peer" (the kernels' own NVLink stores, no collective
call) and compares the reduced gradients; then times each.  Prints one JSON line on rank 0.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29511 tools/dp_check.py [--gaussians 1000000]"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tinysplat_b200 import synthetic  # noqa: E402
from tinysplat_b200.parallel import DataParallelRenderer  # noqa: E402
from tinysplat_b200.rasterizer import GaussianRasterizer, ParamModel  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gaussians", type=int, default=1_000_000)
    ap.add_argument("--steps", type=int, default=20)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    W, H, N = 1920, 1080, args.gaussians
    sc = synthetic.make_scene(N, W, H, seed=0)
    cot = torch.rand(H, W, 3, generator=torch.Generator().manual_seed(7)).to(dev) / (3 * W * H)
    cot_d = torch.rand(H, W, generator=torch.Generator().manual_seed(8)).to(dev) / (W * H)
    names = ["means", "scales", "quats", "opacities", "colors_dc", "colors_rest"]
    out, ms = {}, {}
    for strategy in ("allreduce", "packed", "peer"):
        model = ParamModel(sc, dev, 3)
        rast = GaussianRasterizer(model, None, dev, "fused")
        dp = DataParallelRenderer(rast, model.parameters(), average=True, strategy=strategy)

        def step(i):
            cam = synthetic.make_camera(W, H, yaw_deg=3.0 * rank + 0.1 * i, shift=(0.02 * rank, 0.0, 0.0))
            img, ex = rast(cam, (W, H), 3)
            torch.autograd.backward([img, ex["depth"]], [cot, 0.01 * cot_d])
            dp.reducer.finish()
            return ex

        ex = step(0)
        torch.cuda.synchronize()
        out[strategy] = [getattr(model, k).grad.clone() for k in names] + [ex["xys"].grad.clone()]
        for i in range(3):
            model.zero_grad()
            step(i)
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            model.zero_grad()
            step(i)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms[strategy] = t.item()
        dp.reducer.close()
        if rast.grad_exchange is not None and hasattr(rast.grad_exchange, "check"):
            rast.grad_exchange.check()          # raises if a peer barrier timed out
            sent = rast.grad_exchange.last_bytes_sent
            rast.grad_exchange.close()
        elif rast.grad_exchange is not None:
            sent = rast.grad_exchange.last_bytes_sent
        else:
            sent = None
        ms[strategy + "_bytes_sent_per_rank"] = sent
    report = {"world": world, "gaussians": N, "ms_per_step": ms}
    ok = True
    for strategy in ("packed", "peer"):
        errs = {}
        for k, a, b in zip(names + ["xys"], out["allreduce"], out[strategy]):
            errs[k] = ((a - b).abs().max() / a.abs().max().clamp_min(1e-30)).item()
        worst = torch.tensor([max(errs.values())], device=dev)
        dist.all_reduce(worst, op=dist.ReduceOp.MAX)
        # every rank must hold the same reduced gradient (bit for bit: same shard kernel output everywhere)
        chk = torch.stack([t.double().sum() for t in out[strategy][:6]])
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        report[f"rel_err_{strategy}_vs_allreduce"] = errs
        report[f"{strategy}_worst_over_ranks"] = worst.item()
        report[f"{strategy}_identical_across_ranks"] = bool((hi - lo).abs().max().item() == 0.0)
        ok = ok and worst.item() < 1e-4 and report[f"{strategy}_identical_across_ranks"]
    report["ok"] = bool(ok)
    if rank == 0:
        print(json.dumps(report), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
