"""Times the collectives of the two gradient-exchange strategies in isolation (torchrun, one rank per
GPU): all-reduce of the 236 B/Gaussian gradient span vs all-to-all of packed rows + all-gather of the
shard gradients (coalesced / separate / one flat buffer).  Device time by CUDA events (max over
ranks) and host time of the issuing calls.  Prints one JSON line on rank 0."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tinysplat_b200.parallel import PackedGradExchange  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    N, K = 1_000_000, 16
    ex = PackedGradExchange(average=True)
    Ns = ex.shard_rows(N)
    f32 = dict(device=dev, dtype=torch.float32)
    send = torch.randn(world * Ns, 12, **f32)
    shards = [torch.randn(Ns, K - 1, 3, **f32), torch.randn(Ns, 1, 3, **f32), torch.randn(Ns, 3, **f32),
              torch.randn(Ns, 3, **f32), torch.randn(Ns, 4, **f32), torch.randn(Ns, **f32)]
    flat_shard = torch.randn(Ns * 59, **f32)
    flat_full = torch.empty(world * Ns * 59, **f32)
    span = torch.randn(N * 59, **f32)
    recv = torch.empty(world, Ns, 12, **f32)
    outs = [torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), **f32) for t in shards]

    def a2a():
        dist.all_to_all_single(recv, send)

    def ag_coalesced():
        with dist._coalescing_manager(device=dev, async_ops=False):
            for o, t in zip(outs, shards):
                dist.all_gather_into_tensor(o, t)

    def ag_separate():
        for o, t in zip(outs, shards):
            dist.all_gather_into_tensor(o, t)

    def ag_flat():
        dist.all_gather_into_tensor(flat_full, flat_shard)

    def allreduce():
        dist.all_reduce(span, op=dist.ReduceOp.AVG)

    def exchange_api():
        r = ex.all_to_all_rows(send)
        ex.all_gather_shards(shards)
        return r

    res = {}
    for name, fn in (("all_to_all_rows", a2a), ("all_gather_coalesced", ag_coalesced),
                     ("all_gather_separate", ag_separate), ("all_gather_flat", ag_flat),
                     ("all_reduce_span", allreduce), ("exchange_api(a2a+allgather, allocating)", exchange_api)):
        for _ in range(3):
            fn()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        host = (time.time() - t0) / 10 * 1e3
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 10], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[name] = {"device_ms": round(t.item(), 4), "host_issue_ms": round(host, 4)}
    if rank == 0:
        print(json.dumps({"world": world, "gaussians": N, "shard_rows": Ns, "results": res}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
