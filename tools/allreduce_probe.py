"""All-reduce microbenchmark at the gradient payload size (236 MB fp32), for NCCL algorithm choice."""
import os, sys, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 59_000_000
x = torch.randn(n, device="cuda")
for _ in range(5):
    dist.all_reduce(x, op=dist.ReduceOp.AVG)
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
iters = 20
for _ in range(iters):
    dist.all_reduce(x, op=dist.ReduceOp.AVG)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
if rank == 0:
    gb = n * 4 / 1e9
    print(f"algo={os.environ.get('NCCL_ALGO','auto')} proto={os.environ.get('NCCL_PROTO','auto')} nvls={os.environ.get('NCCL_NVLS_ENABLE','-')} "
          f"bytes={n*4/1e6:.0f}MB  {ms:.3f} ms  algbw={gb/ms*1e3:.0f} GB/s  busbw={gb/ms*1e3*2*(world-1)/world:.0f} GB/s", flush=True)
dist.destroy_process_group()
