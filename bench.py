#!/usr/bin/env python
"""bench.py — headline benchmark of the hot path: Mpix/s forward+backward of the Gaussian
rasterizer (project -> SH -> tile bin/sort -> alpha-blend, and back), one view per GPU per step.

  python bench.py --gpus N --steps K --warmup W            (our arm; torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...  (CPU restatement on the host cores)

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for what each key means.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "Mpix/s fwd+bwd @1M Gaussians/1080p"
UNIT = "Mpix/s"

# name -> (N gaussians, W, H, sh degree, depth loss weight, forward only)
WORKLOADS = {
    # the configuration the metric is quoted on (synthetic stand-in: no dataset on disk)
    "synthetic_1M_1080p": (1_000_000, 1920, 1080, 3, 0.0, False),
    # BASELINE configs[1] (T&T truck stand-in), configs[2], configs[4]
    "synthetic_500k_1080p": (500_000, 1920, 1080, 3, 0.0, False),
    "synthetic_2M_1080p_depthreg": (2_000_000, 1920, 1080, 3, 0.2, False),
    "synthetic_4M_4k_fwd": (4_000_000, 3840, 2160, 3, 0.0, True),
    # dense stand-in for real captures (mean projected radius ~20 px instead of ~6: tile lists in the
    # thousands, several staged batches per tile in the blend kernels, the CTA sort classes of K3)
    "synthetic_dense_1M_1080p": (1_000_000, 1920, 1080, 3, 0.0, False),
    # small case for quick checks
    "synthetic_100k_720p": (100_000, 1280, 720, 3, 0.0, False),
}
SCENE_KW = {"synthetic_dense_1M_1080p": {"mean_radius_px": 20.0}}
DEFAULT_WORKLOAD = "synthetic_1M_1080p"
# extra measurements carried by the default single-GPU line (bench.py with no --workload/--pipeline):
EXTRA_WORKLOADS = ["synthetic_500k_1080p", "synthetic_2M_1080p_depthreg", "synthetic_4M_4k_fwd",
                   "synthetic_dense_1M_1080p"]


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--pipeline", default="fused", choices=["fused", "reference", "unfused4"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--grad-exchange", default="auto", choices=["auto", "allreduce", "packed", "peer"],
                    help="N > 1: how the per-Gaussian gradients are reduced over the ranks' views "
                         "(auto = time every strategy during warm-up, keep the fastest)")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the drop-in / other-workload measurements of the default single-GPU line")
    ap.add_argument("--sustained-s", type=float, default=3.0,
                    help="single GPU: also loop the step for this many seconds (0 = off)")
    ap.add_argument("--cpu-window", type=int, default=8, help="CPU sample: window edge in tiles")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------
def view_for(step: int, rank: int, W: int, H: int):
    """A different (deterministic) camera per rank and step: small yaw + lateral shift."""
    from tinysplat_b200 import synthetic
    k = step * 131 + rank * 17
    yaw = ((k * 37) % 160) / 10.0 - 8.0
    sx = ((k * 53) % 100) / 250.0 - 0.2
    sy = ((k * 29) % 100) / 500.0 - 0.1
    return synthetic.make_camera(W, H, yaw_deg=yaw, shift=(sx, sy, 0.0))


class ClockSampler:
    """Samples SM clock, power and clock-event (throttle) reasons of one GPU through NVML every
    few milliseconds on a background thread, DURING the timed region."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
               0x4: "sw_power_cap", 0x80: "hw_power_brake_slowdown"}

    def __init__(self, device_index: int, period_s: float = 0.004):
        self.period = period_s
        self.samples = []
        self.thread = None
        self.stop_flag = threading.Event()
        self.handle = None
        self.err = None
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            try:
                uuid = str(torch.cuda.get_device_properties(device_index).uuid)
                uuid = uuid if uuid.startswith("GPU-") else "GPU-" + uuid
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception as exc:  # NVML missing: report it, never fail the bench
            self.err = repr(exc)

    def start(self):
        if self.handle is None:
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        nv, h = self.nv, self.handle
        while not self.stop_flag.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                self.samples.append((time.time(), float(sm), int(rs), pw))
            except Exception as exc:
                self.err = repr(exc)
                return
            time.sleep(self.period)

    def stop(self, t0: float, t1: float):
        if self.handle is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [f"nvml unavailable: {self.err}"]}
        self.stop_flag.set()
        self.thread.join(timeout=2)
        win = [x for x in self.samples if t0 <= x[0] <= t1]
        if not win:   # region shorter than one period: take the samples closest to it
            win = sorted(self.samples, key=lambda x: abs(x[0] - 0.5 * (t0 + t1)))[:3]
        reasons = set()
        for _, _, rs, _ in win:
            for bit, name in self.REASONS.items():
                if rs & bit:
                    reasons.add(name)
        sm = [x[1] for x in win]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.sm_max,
                "power_w_max": max((x[3] for x in win), default=None), "samples": len(win),
                "reasons": sorted(reasons)}


# dram__bytes_read.sum + dram__bytes_write.sum per launch, from the `ncu --set full` capture of this
# same command summarised in profiles/r1_ncu_full_all_kernels.txt and, for the blend kernels of the
# final build of round 2, profiles/r2x_ncu_full_all_kernels.txt (default workload, fused pipeline)
NCU_TRAFFIC_SOURCE = ("constant from one `ncu --set full` capture of this command (profiles/), per launch; "
                      "not re-measured in this run")
# sm__inst_issued.avg.pct_of_peak_sustained_active of the same capture: the blend kernels' real limiter
NCU_ISSUE_PCT = {"ts_blend_bwd": 77.6, "ts_blend_fwd": 87.6}
NCU_TRAFFIC = {
    "synthetic_1M_1080p": {
        "ts_blend_bwd": (187.99 + 24.18) * 1e6, "ts_blend_fwd": (62.85 + 22.86) * 1e6,
        "ts_sh_fwd": (228.72 + 24.87) * 1e6, "ts_sh_bwd": (61.03 + 134.51) * 1e6,
        "ts_project_fwd": (62.80 + 23.48) * 1e6, "ts_project_bwd": (95.97 + 23.53) * 1e6,
        "ts_project_sh_bwd": (97.04 + 186.62) * 1e6,
        "ts_bin_emit": (56.31 + 1.16) * 1e6, "ts_bin_sort": (16.08 + 0.0) * 1e6,
    }
}


def algorithmic_bytes(name: str, N: int, M: int, P: int, CH: int, K: int, nb: int) -> float:
    """Algorithmic HBM bytes of one launch of C-ABI entry `name` (per-kernel table).
    N Gaussians, M tile intersections, P pixels, CH blended channels, K stored / nb active SH
    bases."""
    table = {
        "ts_project_fwd": 96 * N,                      # 40 in, 56 out (fused: 44 in, 48 out, + 4 M count atomics)
        "ts_project_bwd": 108 * N,                     # 68 in, 40 out
        "ts_sh_fwd": (24 + 12 * nb) * N,
        "ts_sh_bwd": (24 + 12 * K) * N,
        # K6 + K7 in one kernel: 97 in (means, scales, quats, radii, packed row, logit, mask) + 52 + 12 K out
        "ts_project_sh_bwd": (97 + 52 + 12 * K) * N,
        "ts_bin_count": (76 + 4 * CH) * N + 4 * M,     # pack 48 B record + count atomics
        "ts_bin_scan": 8 * (P // 256),
        "ts_bin_emit": 24 * N + 12 * M,
        "ts_bin_sort": 12 * M,
        "ts_blend_fwd": 52 * M + (8 + 4 * CH) * P,     # id + 48 B record per pair; image, T, n
        "ts_blend_bwd": 52 * M + (12 + 4 * CH) * P + 96 * N,   # + zero & RMW of packed grads
        "ts_blend_unpack_grads": (64 + 24 + 4 * CH) * N,
    }
    return float(table.get(name, 0))


# ---------------------------------------------------------------------------------------------
def host_threads() -> int:
    """The host threads this process may use.  torchrun exports OMP_NUM_THREADS=1 to its workers, so
    the CPU arms set the thread count explicitly instead of inheriting it."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def static_config(workload: str, world: int) -> dict:
    """The part of `config` that names the workload: identical in our arm and in --impl reference."""
    N, W, H, deg, depth_w, fwd_only = WORKLOADS[workload]
    return {"workload": workload, "gaussians": N, "width": W, "height": H, "sh_degree": deg,
            "views_per_step": world, "mode": "fwd" if fwd_only else "fwd+bwd", "depth_loss_weight": depth_w,
            "mean_radius_px": SCENE_KW.get(workload, {}).get("mean_radius_px", 6.0),
            "parallelism": ("single GPU" if world == 1 else
                            (f"replicas only x{world}" if fwd_only else f"dp{world} over cameras, replica per GPU")),
            "l2": "inputs larger than L2: 236 B/Gaussian parameters + 48 B records + image buffers "
                  "> 126 MB; a different camera every step"}


def cpu_sample(workload: str, window_tiles: int, steps: int, warmup: int, threads: int | None = None):
    """Times the CPU oracle (our PyTorch restatement — the reference has no CPU rasterizer) on a
    bounded sample of the workload: the central window of `window_tiles`^2 tiles, with the
    Gaussians that can reach it.  Returns (Mpix/s, description, cores, per-step seconds)."""
    import oracle  # the checker, used here only as the reported CPU baseline
    from tinysplat_b200 import synthetic
    N, W, H, deg, depth_w, fwd_only = WORKLOADS[workload]
    torch.set_num_threads(threads or host_threads())
    cores = torch.get_num_threads()
    cam = synthetic.make_camera(W, H)
    sc = synthetic.make_scene(N, W, H, seed=0, **SCENE_KW.get(workload, {}))
    tbx, tby = (W + 15) // 16, (H + 15) // 16
    tx0, ty0 = (tbx - window_tiles) // 2, (tby - window_tiles) // 2
    win = (tx0, ty0, tx0 + window_tiles, ty0 + window_tiles)
    # sample preparation (untimed): keep Gaussians whose 3-sigma box can reach the window
    with torch.no_grad():
        q = sc["quats"]
        V, Pm = cam.view_matrix, cam.proj_matrix
        xys, _, radii, _, _, _ = oracle.project_gaussians(
            sc["means"], torch.exp(sc["scales"]), 1.0, q / q.norm(dim=-1, keepdim=True), V[:3], Pm @ V,
            cam.f_x, cam.f_y, W / 2, H / 2, H, W, (tbx, tby, 1))
        r = radii.float() + 16
        keep = (radii > 0) & (xys[:, 0] + r >= win[0] * 16) & (xys[:, 0] - r <= win[2] * 16) & \
               (xys[:, 1] + r >= win[1] * 16) & (xys[:, 1] - r <= win[3] * 16)
    sub = {k: (v[keep].clone() if k != "background" else v) for k, v in sc.items()}
    n_sub = int(keep.sum())
    names = ["means", "scales", "quats", "opacities", "colors_dc", "colors_rest"]
    pix = (window_tiles * 16) ** 2
    g = torch.Generator().manual_seed(1)
    gt = torch.rand(window_tiles * 16, window_tiles * 16, 3, generator=g)
    times = []
    for it in range(warmup + steps):
        p = {k: (v.clone().requires_grad_(not fwd_only) if k in names else v) for k, v in sub.items()}
        t0 = time.perf_counter()
        img, ex = oracle.render_reference_adapter(p, cam.view_matrix, cam.proj_matrix, cam.f_x, cam.f_y,
                                                  (W, H), deg, tile_window=win)
        if not fwd_only:
            loss = (img - gt).abs().mean()
            if depth_w:
                loss = loss + depth_w * ex["depth"].abs().mean()
            loss.backward()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    sec = statistics.mean(times)
    desc = (f"central {window_tiles}x{window_tiles}-tile window ({window_tiles * 16}^2 px) of {workload}, "
            f"{n_sub} Gaussians reaching it, full adapter op sequence "
            f"({'fwd' if fwd_only else 'fwd+bwd'}), fp32 torch CPU, {cores} threads, {steps} steps")
    return pix / sec / 1e6, desc, cores, sec


def run_reference(args):
    """--impl reference: the reference has no CPU (or any in-tree) implementation of this path —
    its arithmetic is the absent gsplat package — so this arm times our CPU restatement (the
    oracle, kind "port") on the host cores, on a bounded sample of the same workload: each step
    renders the central 8x8-tile window fwd+bwd (about 0.3 s), K steps after W warm-up steps."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    val, desc, cores, sec = cpu_sample(args.workload, args.cpu_window, steps, warmup, threads=host_threads())
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": static_config(args.workload, world),
        "note": "CPU restatement (oracle) on a bounded sample; host cores only; the reference has no CPU path",
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc,
                         "host_cpu_count": os.cpu_count(), "omp_num_threads_env": os.environ.get("OMP_NUM_THREADS")},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
class Arm:
    """One (workload, pipeline) instance on this rank: the parameter replica, the adapter, the
    gradient exchange and the two kinds of step the bench times."""

    def __init__(self, workload: str, pipeline: str, dev, rank: int, world: int):
        from tinysplat_b200 import synthetic
        from tinysplat_b200.parallel import GradientAllReducer
        from tinysplat_b200.rasterizer import GaussianRasterizer, ParamModel
        self.workload, self.pipeline, self.dev, self.rank, self.world = workload, pipeline, dev, rank, world
        self.N, self.W, self.H, self.deg, self.depth_w, self.fwd_only = WORKLOADS[workload]
        self.P = self.W * self.H
        sc = synthetic.make_scene(self.N, self.W, self.H, seed=0, **SCENE_KW.get(workload, {}))
        self.model = ParamModel(sc, dev, self.deg, requires_grad=not self.fwd_only)   # same replica on every rank
        self.rast = GaussianRasterizer(self.model, None, dev, pipeline)
        self.reducer = None if self.fwd_only else GradientAllReducer(self.model.parameters(), average=True, overlap=True)
        self.exchange = {"name": "allreduce" if (world > 1 and not self.fwd_only) else None}
        # fixed cotangents for the device-resident arm ("fixed cotangent v_out ~ U(0,1)
        # so backward cost is content-independent"); the loss arithmetic is not part of the hot path
        H, W, P = self.H, self.W, self.P
        self.cot_img = torch.rand(H, W, 3, generator=torch.Generator().manual_seed(7)).to(dev) / (3 * P)
        self.cot_depth = (self.depth_w * torch.rand(H, W, generator=torch.Generator().manual_seed(8)).to(dev) / P) \
            if self.depth_w else None

    def set_exchange(self, name: str):
        """allreduce: NCCL all-reduce of the finished gradients (one flat 236 B/Gaussian span);
        packed: NCCL all-to-all of blend-backward's packed rows + shard backward + all-gather;
        peer: the same shard backward with both transfers done by the kernels themselves over
        NVLink peer memory (tinysplat_b200.parallel).  Fused pipeline only."""
        from tinysplat_b200 import parallel
        if self.reducer is not None:
            self.reducer.close()
        old = getattr(self.rast, "grad_exchange", None)
        if old is not None and hasattr(old, "close"):
            old.close()
        if name == "allreduce":
            self.rast.grad_exchange = None
            self.reducer = parallel.GradientAllReducer(self.model.parameters(), average=True, overlap=True)
        else:
            cls = parallel.PackedGradExchange if name == "packed" else parallel.PeerGradExchange
            self.rast.grad_exchange = cls(average=True)
            self.reducer = parallel.GradientAllReducer([], average=True, overlap=False)
        self.exchange["name"] = name

    def path_step(self, i: int):
        """`value`: the hot path alone — adapter forward, backward from fixed cotangents."""
        cam = view_for(i, self.rank, self.W, self.H)
        if self.fwd_only:
            with torch.no_grad():
                self.rast(cam, (self.W, self.H), self.deg)
            return
        img, ex = self.rast(cam, (self.W, self.H), self.deg)
        if self.cot_depth is not None:
            torch.autograd.backward([img, ex["depth"]], [self.cot_img, self.cot_depth])
        else:
            img.backward(self.cot_img)
        self.reducer.finish()
        self.model.zero_grad()

    def one_step(self, i: int, gt: torch.Tensor):
        """`e2e`: a training step as a user writes it — render, L1 (+depth) loss, backward.  gt: the uint8
        image as loaded from disk (tinysplat_b200.loss.l1_loss) or the float32 tensor the reference uploads."""
        from tinysplat_b200.loss import l1_loss
        cam = view_for(i, self.rank, self.W, self.H)
        if self.fwd_only:
            with torch.no_grad():
                img, ex = self.rast(cam, (self.W, self.H), self.deg)
            return img.mean()
        img, ex = self.rast(cam, (self.W, self.H), self.deg)
        if gt.dtype == torch.uint8:
            loss = l1_loss(img, gt)                      # fused: value + gradient in one pass, u8 / 255 in the kernel
        else:
            loss = (img - gt).abs().mean()               # the reference's expression [REF scripts/train.py:59]
        if self.depth_w:
            loss = loss + self.depth_w * ex["depth"].abs().mean()
        loss.backward()
        self.reducer.finish()
        self.model.zero_grad()
        return loss.detach()

    def close(self):
        if self.reducer is not None:
            self.reducer.close()
        ge = getattr(self.rast, "grad_exchange", None)
        if ge is not None and hasattr(ge, "close"):
            ge.close()
        self.model = self.rast = self.reducer = None
        torch.cuda.empty_cache()


def _barrier(world):
    import torch.distributed as dist
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def timed(fn, steps: int, first_index: int, world: int, dev):
    """K steps bracketed by barrier + synchronize, CUDA events, max over ranks.  -> (ms, t0, t1)."""
    import torch.distributed as dist
    _barrier(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    for i in range(steps):
        fn(first_index + i)
    e1.record()
    _barrier(world)
    t1 = time.time()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return ms.item(), t0, t1


def measure_value(arm: Arm, K: int, Wm: int, local: int, sample_clocks: bool):
    """Device-resident arm.  -> dict(ms_step, value, prof, launches, clocks, M, max_per_tile)."""
    from tinysplat_b200 import _lib, rasterize as rz
    for i in range(Wm):
        arm.path_step(i)
    # per-kernel table: a pass of its own with CUDA events around EVERY C-ABI call (two event records per
    # call cost ~5 % of the step, so it is kept out of the timed region)
    Kp = min(K, 10)
    _lib.profile_start()
    for i in range(Kp):
        arm.path_step(Wm + K + i)
    prof_all = _lib.profile_stop()
    sampler = ClockSampler(local) if sample_clocks else None
    if sampler:
        sampler.start()
    n0 = _lib.launch_count()
    # timed region: only the dominant kernel is timed live (the roofline line's `achieved`)
    dominant = "ts_blend_fwd" if arm.fwd_only else "ts_blend_bwd"
    _lib.profile_start(only=(dominant,))
    ms_total, t0, t1 = timed(arm.path_step, K, Wm, arm.world, arm.dev)
    prof_dom = _lib.profile_stop()
    launches = _lib.launch_count() - n0
    clocks = sampler.stop(t0, t1) if sampler else None
    # the table: every kernel scaled to K steps from the profiling pass, the dominant one from the timed region
    prof = {name: [sum(v) / len(v)] * max(round(len(v) * K / Kp), 1) for name, v in prof_all.items()}
    prof.update(prof_dom)
    ms_step = ms_total / K
    return {"ms_step": ms_step, "ms_total": ms_total, "value": arm.world * arm.P / (ms_step * 1e-3) / 1e6,
            "prof": prof, "launches": int(launches), "clocks": clocks,
            "M": rz.last_stats["num_intersects"], "max_per_tile": rz.last_stats["max_per_tile"]}


def measure_e2e(arm: Arm, K: int, gt_u8: bool = True):
    """End-to-end arm: host buffers in, loss out.  Every step copies ITS target image (pinned host ->
    device) and reads its loss back.  Like a training data loader, step i+1's image is prefetched on a
    copy stream while step i renders; all copies happen inside the timed region.  gt_u8: the target
    travels as the uint8 image it is stored as (6.2 MB at 1080p) and the fused L1 loss forms u8 / 255;
    otherwise as the float32 tensor the reference uploads (24.9 MB) with the loss written in torch."""
    dev, H, W = arm.dev, arm.H, arm.W
    g = torch.Generator().manual_seed(100 + arm.rank)
    if gt_u8:
        gt_host = torch.randint(0, 256, (H, W, 3), generator=g, dtype=torch.uint8).pin_memory()
    else:
        gt_host = torch.rand(H, W, 3, generator=g).pin_memory()
    copy_stream = torch.cuda.Stream(device=dev)
    gt_bufs = [torch.empty(H, W, 3, device=dev, dtype=gt_host.dtype) for _ in range(2)]
    copied = [torch.cuda.Event(), torch.cuda.Event()]      # H2D of buffer k finished
    consumed = [None, None]                                 # compute that read buffer k finished
    host_loss = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_ready = [None, None]
    losses = []
    issued = set()

    def issue_copy(i):
        if i in issued:
            return
        issued.add(i)
        k = i % 2
        if consumed[k] is not None:
            copy_stream.wait_event(consumed[k])              # do not overwrite a buffer still being read
        with torch.cuda.stream(copy_stream):
            gt_bufs[k].copy_(gt_host, non_blocking=True)
            copied[k].record(copy_stream)

    def collect(k):
        if loss_ready[k] is not None:
            loss_ready[k].synchronize()
            losses.append(float(host_loss[k]))
            loss_ready[k] = None

    def e2e_step(i):
        k = i % 2
        issue_copy(i)
        issue_copy(i + 1)                                    # H2D of the next step's target image
        torch.cuda.current_stream().wait_event(copied[k])
        loss = arm.one_step(i, gt_bufs[k])
        consumed[k] = torch.cuda.Event()
        consumed[k].record()
        collect(k)                                           # (buffer k's previous loss, two steps ago)
        host_loss[k].copy_(loss, non_blocking=True)          # D2H of this step's result ...
        loss_ready[k] = torch.cuda.Event()
        loss_ready[k].record()
        collect(1 - k)                                       # ... read on the host one step later

    for i in range(3):
        e2e_step(1000 + i)
    # K steps, timed twice; the faster repetition is reported and both are listed (a host-side hiccup —
    # page faults of the pinned buffers, a scheduler tick — can cost 10 % of a 20 ms region)
    reps = []
    for rep in range(2):
        ms_rep, _, _ = timed(e2e_step, K, 2000 + rep * K, arm.world, dev)   # ends with a device sync: every loss has landed
        reps.append(ms_rep)
    ms_e2e = min(reps)
    collect(0)
    collect(1)
    return {"value": arm.world * arm.P / (ms_e2e / K * 1e-3) / 1e6, "unit": UNIT,
            "ms_per_step_repetitions": [r / K for r in reps],
            "h2d_bytes_per_step": gt_host.numel() * gt_host.element_size() + 2 * 16 * 4,   # target image + view/proj matrices
            "target": "uint8 image, fused L1 loss (tinysplat_b200.loss)" if gt_u8 else "float32 image, torch L1 loss",
            "d2h_bytes_per_step": 4 + 16,                                 # loss scalar + binning stats (4 x int32)
            "ms_per_step": ms_e2e / K}


def measure_train_step(dev, steps: int = 15):
    """The whole training step of scripts/train.py at the default workload, built from this repo's pieces:
    fused adapter, 0.8 L1 + 0.2 (1 - SSIM) with the fused losses, backward, FusedAdam over the six
    parameter groups [REF scripts/train.py:54-100; model_gaussian.py:112-120].  Device-timed."""
    from tinysplat_b200 import synthetic
    from tinysplat_b200.loss import l1_loss
    from tinysplat_b200.optim import FusedAdam
    from tinysplat_b200.rasterizer import GaussianRasterizer, ParamModel
    from tinysplat_b200.ssim import SSIM
    N, W, H, deg, _, _ = WORKLOADS[DEFAULT_WORKLOAD]
    names = ["means", "scales", "quats", "opacities", "colors_dc", "colors_rest"]
    lrs = dict(means=0.00016, colors_dc=0.0025, colors_rest=0.000125, scales=0.005, quats=0.001, opacities=0.05)
    model = ParamModel(synthetic.make_scene(N, W, H, seed=0), dev, deg)
    params = {k: torch.nn.Parameter(getattr(model, k).detach()) for k in names}
    for k, p in params.items():
        setattr(model, k, p)
    opt = FusedAdam([{"params": [params[k]], "lr": lrs[k], "name": k} for k in names])
    ssim = SSIM(data_range=1.0, channel=3)
    rast = GaussianRasterizer(model, None, dev, "fused")
    gt_u8 = torch.randint(0, 256, (H, W, 3), generator=torch.Generator().manual_seed(3), dtype=torch.uint8).to(dev)
    gt_chw = (gt_u8.float() / 255).permute(2, 0, 1)[None]

    def step(i):
        img, _ = rast(view_for(i, 0, W, H), None, deg)
        loss = 0.8 * l1_loss(img, gt_u8) + 0.2 * (1 - ssim(img.permute(2, 0, 1)[None], gt_chw))
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)

    for i in range(4):
        step(i)
    ms, _, _ = timed(step, steps, 100, 1, dev)
    ms /= steps
    del model, params, opt, rast
    torch.cuda.empty_cache()
    return {"what": "render (fused adapter) + 0.8 L1 + 0.2 (1 - SSIM) (fused losses) + backward + FusedAdam, "
                    "1M Gaussians / 1080p, device-resident ground truth",
            "ms_per_step": ms, "value": W * H / (ms * 1e-3) / 1e6, "unit": UNIT, "steps": steps}


def measure_sustained(arm: Arm, seconds: float, local: int):
    """The device-resident step looped for >= `seconds` of wall time with the NVML clock / power
    record: the burst-clock figure of the K-step region next to what the GPU holds under load."""
    sampler = ClockSampler(local, period_s=0.02)
    sampler.start()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    n = 0
    while True:
        for _ in range(50):
            arm.path_step(5000 + n)
            n += 1
        if time.time() - t0 >= seconds:       # the host runs at most one step ahead (mid-step stats read)
            break
    e1.record()
    torch.cuda.synchronize()
    t1 = time.time()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop(t0, t1)
    sm = sorted(x[1] for x in sampler.samples if t0 <= x[0] <= t1)
    if sm:
        clocks["sm_mhz_min"] = sm[0]
        clocks["sm_mhz_p10"] = sm[len(sm) // 10]
    return {"seconds": ms * 1e-3, "steps": n, "ms_per_step": ms / n, "value": arm.P / (ms / n * 1e-3) / 1e6,
            "unit": UNIT, "clocks": clocks}


def kernel_table(prof: dict, ms_total: float, K: int, N, M, P, CH, Kb, nb):
    kern = {}
    for name, ms_list in prof.items():
        tot = sum(ms_list)
        per = tot / len(ms_list)
        by = algorithmic_bytes(name, N, M, P, CH, Kb, nb)
        kern[name] = {"launches_per_step": len(ms_list) / K, "ms_per_launch": per, "ms_per_step": tot / K,
                      "share": tot / ms_total, "algorithmic_bytes": by,
                      "gbs": by / (per * 1e-3) / 1e9 if per > 0 else None}
    return kern


def run_ours(args):
    import torch.distributed as dist
    from tinysplat_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (our arm) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()
    K, Wm = max(1, args.steps), max(3, args.warmup)

    arm = Arm(args.workload, args.pipeline, dev, rank, world)
    N, W, H, deg, depth_w, fwd_only = WORKLOADS[args.workload]
    P = W * H
    probe = {}

    # ---- N > 1: pick the gradient-exchange strategy (untimed, before the warm-up proper) ------
    if world > 1 and not fwd_only:
        choice = args.grad_exchange if args.pipeline == "fused" else "allreduce"
        if choice == "auto":
            for name in ("allreduce", "packed", "peer"):
                try:
                    arm.set_exchange(name)
                    for i in range(3):
                        arm.path_step(i)
                    probe[name] = timed(arm.path_step, 5, 3, world, dev)[0] / 5   # max over ranks: same everywhere
                except Exception as e:                           # deterministic errors hit every rank alike
                    probe[name] = float("inf")
                    probe[name + "_error"] = f"{type(e).__name__}: {e}"[:200]
                    arm.model.zero_grad()
            # every rank must take the same decision: rank 0's (the timings are max-reduced, but a
            # failure on one rank only must not split the job)
            best = min((n for n in ("allreduce", "packed", "peer")), key=lambda n: probe[n])
            pick = torch.tensor([("allreduce", "packed", "peer").index(best)], device=dev)
            dist.broadcast(pick, 0)
            choice = ("allreduce", "packed", "peer")[int(pick.item())]
        arm.set_exchange(choice)

    main = measure_value(arm, K, Wm, local, sample_clocks=(rank == 0))
    e2e = measure_e2e(arm, K)
    if not arm.fwd_only:
        # the same step fed the way the reference feeds it (float32 image uploaded, loss written in torch)
        ref_style = measure_e2e(arm, K, gt_u8=False)
        e2e["float32_target_torch_loss"] = {k: ref_style[k] for k in ("value", "ms_per_step", "h2d_bytes_per_step")}
    sustained = None
    if world == 1 and args.sustained_s > 0:
        sustained = measure_sustained(arm, args.sustained_s, local)
    exch_name = arm.exchange["name"]
    exch_bytes = getattr(arm.rast.grad_exchange, "last_bytes_sent", None) if arm.rast.grad_exchange is not None else None
    n_coll = arm.reducer.last_num_collectives if arm.reducer else 0
    allreduce_bytes = arm.reducer.payload_bytes() if (arm.reducer and world > 1) else 0
    arm.close()

    # ---- the rest of the default single-GPU line: drop-in surface, other workloads ------------
    default_line = (world == 1 and args.workload == DEFAULT_WORKLOAD and args.pipeline == "fused"
                    and not args.no_extras)
    dropin, workloads, train_step = None, None, None
    if default_line:
        try:
            train_step = measure_train_step(dev)
        except Exception as exc:
            train_step = {"error": repr(exc)[:300]}
        try:
            a = Arm(DEFAULT_WORKLOAD, "reference", dev, rank, world)
            mv = measure_value(a, K, Wm, local, sample_clocks=False)
            me = measure_e2e(a, K, gt_u8=False)          # fed the reference's way: float32 image, torch loss
            kt = kernel_table(mv["prof"], mv["ms_total"], K, N, mv["M"], P, 3, 16, (deg + 1) ** 2)
            dropin = {"what": "the reference adapter's own op sequence through the five gsplat symbols "
                              "(project_gaussians, sh.spherical_harmonics, rasterize_gaussians x2 + torch glue), "
                              "no change to the reference [REF rasterize.py:26-62]",
                      "value": mv["value"], "unit": UNIT, "ms_per_step": mv["ms_step"], "e2e": me["value"],
                      "e2e_ms_per_step": me["ms_per_step"], "gpu_launches": mv["launches"],
                      "intersections_M": mv["M"], "raster_passes_per_render": 2,
                      "ts_kernels_ms_per_step": sum(v["ms_per_step"] for v in kt.values()),
                      "kernels_ms_per_step": {k: round(v["ms_per_step"], 4) for k, v in kt.items()}}
            a.close()
        except Exception as exc:
            dropin = {"error": repr(exc)[:300]}
        workloads = {}
        for wname in EXTRA_WORKLOADS:
            try:
                a = Arm(wname, "fused", dev, rank, world)
                mv = measure_value(a, 10, 3, local, sample_clocks=False)
                kt = kernel_table(mv["prof"], mv["ms_total"], 10, a.N, mv["M"], a.P, 4, 16, (a.deg + 1) ** 2)
                workloads[wname] = {"value": mv["value"], "unit": UNIT, "ms_per_step": mv["ms_step"], "steps": 10,
                                    "warmup": 3, "mode": "fwd" if a.fwd_only else "fwd+bwd",
                                    "gaussians": a.N, "width": a.W, "height": a.H, "depth_loss_weight": a.depth_w,
                                    "intersections_M": mv["M"], "max_per_tile": mv["max_per_tile"],
                                    "kernels_ms_per_step": {k: round(v["ms_per_step"], 4) for k, v in kt.items()}}
                a.close()
            except Exception as exc:
                workloads[wname] = {"error": repr(exc)[:300]}

    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- per-kernel breakdown and roofline of the dominant kernel ----------------------------
    CH = 3 if args.pipeline == "reference" else 4
    Kb = 16 if deg <= 3 else 25
    nb = (deg + 1) ** 2
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    M = main["M"]
    kern = kernel_table(main["prof"], main["ms_total"], K, N, M, P, CH, Kb, nb)
    top = max(kern, key=lambda k: kern[k]["ms_per_step"]) if kern else None
    roof = None
    if top:
        a = kern[top]["gbs"]
        traffic = NCU_TRAFFIC.get(args.workload, {}).get(top) if args.pipeline == "fused" else None
        roof = {"kernel": top, "bound": "hbm", "achieved": a, "peak": peak, "unit": "GB/s",
                "frac": a / peak, "traffic": traffic, "peak_source": peak_src,
                "traffic_source": NCU_TRAFFIC_SOURCE if traffic else None,
                "note": "blend kernels are fp32-issue/atomic bound, not HBM bound (DESIGN.md); "
                        "streaming kernels listed in `kernels`",
                # what does bound it (one `ncu --set full` capture of this command, profiles/r2x_ncu_full_all_kernels.txt)
                "issue_slots_busy_pct_ncu": NCU_ISSUE_PCT.get(top) if args.pipeline == "fused" else None}
    # the best streaming (genuinely HBM-bound) kernel, for a roofline fraction that means something
    stream_names = [k for k in ("ts_sh_bwd", "ts_sh_fwd", "ts_project_sh_bwd", "ts_project_bwd", "ts_project_fwd") if k in kern]
    best_stream = max(stream_names, key=lambda k: kern[k]["gbs"] or 0) if stream_names else None
    roof_stream = None
    if best_stream:
        roof_stream = {"kernel": best_stream, "bound": "hbm", "achieved": kern[best_stream]["gbs"], "peak": peak,
                       "unit": "GB/s", "frac": kern[best_stream]["gbs"] / peak,
                       "traffic": NCU_TRAFFIC.get(args.workload, {}).get(best_stream)}
    cfg = static_config(args.workload, world)
    run_info = {"pipeline": args.pipeline, "intersections_M": M, "max_per_tile": main["max_per_tile"],
                "raster_passes_per_render": 2 if args.pipeline == "reference" else 1,
                "value_step": "adapter forward (RGB+depth) + backward from fixed cotangents (SURVEY 8d)"
                              + (" + gradient exchange" if world > 1 and not fwd_only else ""),
                "e2e_step": "H2D target image + camera, adapter forward, L1 loss, backward, loss.item()",
                "grad_exchange": exch_name,
                "grad_exchange_what": {None: None,
                                       "allreduce": "NCCL all-reduce of the finished gradients (one flat span)",
                                       "packed": "NCCL all-to-all of packed rows + shard backward + NCCL all-gather",
                                       "peer": "rank-structured exchange by the kernels themselves over NVLink peer "
                                               "memory: packed geometry rows to the owner, colour cotangents to "
                                               "every rank, shard projection-backward stores into every rank's "
                                               "gradient buffer; no NCCL call in the step"}[exch_name],
                "probe_allreduce_ms": probe.get("allreduce"), "probe_packed_ms": probe.get("packed"),
                "probe_peer_ms": probe.get("peer"),
                "probe_errors": {k: v for k, v in probe.items() if k.endswith("_error")} or None,
                "grad_allreduce_collectives_per_step": n_coll, "grad_allreduce_bytes": allreduce_bytes,
                "grad_exchange_bytes_sent_per_rank": exch_bytes}
    for k in ("probe_allreduce_ms", "probe_packed_ms", "probe_peer_ms"):
        if run_info[k] == float("inf"):
            run_info[k] = None
    line = {
        "metric": METRIC, "value": main["value"], "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": main["ms_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": cfg,
        "e2e": e2e,
        "gpu_launches": main["launches"],
        "clocks": main["clocks"],
        "roofline": roof,
        "roofline_streaming": roof_stream,
        "run_info": run_info,
        # flat copies of the exchange facts (nested dicts were dropped from the driver's record in round 1)
        "grad_exchange": exch_name, "probe_allreduce_ms": run_info["probe_allreduce_ms"],
        "probe_packed_ms": run_info["probe_packed_ms"], "probe_peer_ms": run_info["probe_peer_ms"],
        "kernels": kern,
        "kernels_source": "CUDA events around each C-ABI call: the dominant kernel (roofline.kernel) inside the timed "
                          "region, the others in a profiling pass of the same step right before it (two event records "
                          "per call cost ~5 % of the step)",
    }
    if sustained is not None:
        sustained["burst_value"] = main["value"]
        sustained["ratio_to_burst"] = sustained["value"] / main["value"]
        line["sustained"] = sustained
        line["sustained_value"] = sustained["value"]
    if dropin is not None:
        line["dropin"] = dropin
        line["dropin_value"] = dropin.get("value")
        line["dropin_e2e"] = dropin.get("e2e")
    if workloads is not None:
        line["workloads"] = workloads
    if train_step is not None:
        line["train_step"] = train_step
    if world == 1 and not args.no_cpu_baseline:
        try:
            v, desc, cores, sec = cpu_sample(args.workload, args.cpu_window, 8, 2)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc,
                                    "host_cpu_count": os.cpu_count(), "ms_per_sample_step": sec * 1e3}
        except Exception as exc:  # the baseline is reporting, never a reason to lose the GPU number
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": None, "kind": "port",
                                    "sample": f"failed: {exc!r}"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
