"""Timeline of one peer-memory gradient exchange per rank (torchrun, one rank per GPU): timing events
recorded by ts_dp_exchange_peer between its launches on the three streams (csrc/peer.cu).  Prints, for
rank 0 and as the max over ranks, the ms since the start of the exchange at which each piece was pushed,
had landed from every rank, had its SH gradient / shard projection gradient done.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
      --master-port 29521 tools/exchange_timeline.py [--chunks 1]"""
import argparse
import ctypes as C
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tinysplat_b200 import _lib, synthetic  # noqa: E402
from tinysplat_b200.parallel import DataParallelRenderer  # noqa: E402
from tinysplat_b200.rasterizer import GaussianRasterizer, ParamModel  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gaussians", type=int, default=1_000_000)
    ap.add_argument("--chunks", type=int, nargs="+", default=[1, 2])
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    W, H, N = 1920, 1080, args.gaussians
    sc = synthetic.make_scene(N, W, H, seed=0)
    cot = torch.rand(H, W, 3, generator=torch.Generator().manual_seed(7)).to(dev) / (3 * W * H)
    report = {"world": world, "gaussians": N}
    for chunks in args.chunks:
        model = ParamModel(sc, dev, 3)
        rast = GaussianRasterizer(model, None, dev, "fused")
        dp = DataParallelRenderer(rast, model.parameters(), average=True, strategy="peer")
        rast.grad_exchange.n_chunks = chunks

        def step(i):
            cam = synthetic.make_camera(W, H, yaw_deg=3.0 * rank + 0.1 * i, shift=(0.02 * rank, 0.0, 0.0))
            img, ex = rast(cam, (W, H), 3)
            img.backward(cot)
            dp.reducer.finish()

        for i in range(5):
            model.zero_grad()
            step(i)
        lib.ts_dp_exchange_timeline(1)
        rows = []
        for i in range(5, 10):
            model.zero_grad()
            dist.barrier()
            torch.cuda.synchronize()
            step(i)
            torch.cuda.synchronize()
            buf = C.create_string_buffer(4096)
            lib.ts_dp_exchange_timeline_read(buf, 4096)
            marks = [ln.split() for ln in buf.value.decode().strip().splitlines()]
            rows.append({k: float(v) for k, v in marks})
        lib.ts_dp_exchange_timeline(0)
        labels = list(rows[0].keys())
        mine = torch.tensor([[r[k] for k in labels] for r in rows], device=dev).median(dim=0).values
        worst = mine.clone()
        dist.all_reduce(worst, op=dist.ReduceOp.MAX)
        report[f"chunks{chunks}"] = {"labels": labels, "rank0_ms": [round(x, 4) for x in mine.tolist()],
                                     "max_over_ranks_ms": [round(x, 4) for x in worst.tolist()]}
        rast.grad_exchange.check()
        rast.grad_exchange.close()
    if rank == 0:
        print(json.dumps(report), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
