"""Single-GPU microbenchmark of the data-parallel exchange kernels with synthetic inputs of the 8-rank
shape (no peers needed: the pointer tables all point into local memory): ts_dp_push (world copies of
the colour rows), ts_sh_bwd_views_rgb (SH gradient of ALL Gaussians from `views` colour-cotangent
rows), ts_project_bwd_views_peer (projection-backward of a 1/world shard over all views, stored to
`world` destinations).  Prints one JSON line; run it under ncu for the per-kernel counters.

  python tools/bench_exchange_kernels.py [--gaussians 1000000] [--views 8]"""
import argparse
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tinysplat_b200 import _lib, synthetic  # noqa: E402
from tinysplat_b200.parallel import PeerLayout  # noqa: E402


def timed(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush = torch.empty(64 * 1024 * 1024, device="cuda", dtype=torch.float32)
    tot = 0.0
    for _ in range(iters):
        flush.zero_()                      # 256 MB written: evicts the 126 MB L2
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gaussians", type=int, default=1_000_000)
    ap.add_argument("--views", type=int, default=8)
    args = ap.parse_args()
    _lib.load()
    dev = torch.device("cuda:0")
    N, V, K, W, H = args.gaussians, args.views, 16, 1920, 1080
    sc = synthetic.make_scene(N, W, H, seed=0)
    f32 = dict(device=dev, dtype=torch.float32)
    means, scales, quats = (sc[k].to(dev).float().contiguous() for k in ("means", "scales", "quats"))
    logits = sc["opacities"].to(dev).float().reshape(-1).contiguous()
    Ns = PeerLayout.shard_rows(N, V)
    Npad = V * Ns
    cams = torch.zeros(V, 32, **f32)
    for v in range(V):
        cam = synthetic.make_camera(W, H, yaw_deg=3.0 * v)
        view, full = cam.view_matrix.float(), (cam.proj_matrix @ cam.view_matrix).float()
        cams[v] = torch.cat([view[:3].reshape(-1), full.reshape(-1), torch.tensor([cam.f_x, cam.f_y, 0.0, 0.0])]).to(dev)
    st = _lib.stream_ptr(dev)
    out = {"gaussians": N, "views": V}
    # -- SH gradient of all Gaussians from V colour-cotangent rows ------------------------------
    rgb = torch.randn(V, Npad, 3, **f32) * 1e-3
    v_dc, v_rest = torch.empty(N, 3, **f32), torch.empty(N, K - 1, 3, **f32)
    ms = timed(lambda: _lib.call("ts_sh_bwd_views_rgb", V, N, 3, K, _lib.ptr(means), _lib.ptr(cams), _lib.ptr(rgb),
                                 Npad * 3, 1.0 / V, _lib.ptr(v_dc), _lib.ptr(v_rest), st))
    nbytes = N * (12 + 12 * V + 12 * K)
    out["ts_sh_bwd_views_rgb"] = {"ms": ms, "algorithmic_bytes": nbytes, "gbs": nbytes / ms / 1e6}
    # -- projection-backward of one shard over V views, stored to V destinations ------------------
    ns = min(Ns, N)
    geo = torch.randn(V, Ns, 8, **f32) * 1e-3
    dsts = [[torch.empty(ns, w, **f32) for _ in range(V)] for w in (3, 3, 4, 1)]
    tabs = [(C.c_void_p * V)(*[t.data_ptr() for t in d]) for d in dsts]
    flags = _lib.PROJ_LOG_SCALES | _lib.PROJ_RAW_QUATS
    ms = timed(lambda: _lib.call("ts_project_bwd_views_peer", V, ns, _lib.ptr(means), _lib.ptr(scales), 1.0,
                                 _lib.ptr(quats), _lib.ptr(cams), H, W, flags, _lib.ptr(geo), Ns * 8, _lib.ptr(logits),
                                 1.0 / V, V, 1, tabs[0], tabs[1], tabs[2], tabs[3], st))
    nbytes = ns * (44 + 32 * V + 44 * V)
    out["ts_project_bwd_views_peer"] = {"ms": ms, "rows": ns, "algorithmic_bytes": nbytes, "gbs": nbytes / ms / 1e6}
    # -- push: geometry rows to the owner, colour rows to V destinations (all local here) -----------
    radii = torch.ones(N, device=dev, dtype=torch.int32)
    mask = torch.full((N,), 7, device=dev, dtype=torch.uint8)
    recs, grads = torch.randn(N, 12, **f32), torch.randn(N, 12, **f32)
    geo_d = [torch.empty(V, Ns, 8, **f32) for _ in range(V)]
    rgb_d = [torch.empty(V, Npad, 3, **f32) for _ in range(V)]
    cam_d = [torch.empty(V, 32, **f32) for _ in range(V)]
    tg, tr, tc = ((C.c_void_p * V)(*[t.data_ptr() for t in d]) for d in (geo_d, rgb_d, cam_d))
    v_xys = torch.empty(N, 2, **f32)
    ms = timed(lambda: _lib.call("ts_dp_push", N, Ns, Npad, V, 0, _lib.ptr(radii), _lib.ptr(mask), _lib.ptr(recs),
                                 _lib.ptr(grads), _lib.ptr(cams[0]), tg, tr, tc, _lib.ptr(v_xys), st))
    nbytes = N * (4 + 1 + 16 + 48 + 8 + 32 + 12 * V)
    out["ts_dp_push_local"] = {"ms": ms, "algorithmic_bytes": nbytes, "gbs": nbytes / ms / 1e6}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
