"""One autograd node for the whole raster adapter  [REF tinysplat/splatting/rasterize.py:26-62].

The reference's host code issues ~25 small torch ops around the three gsplat calls (exp,
normalise, sigmoid, view-direction arithmetic, +0.5/clamp, cat, repeat ...), each with its own
backward kernels; on a B200 the step is then bound by CPU launch overhead, not by the GPU.
Here every one of those activations is folded into the sm_100a kernels (flags of the C ABI),
SH colour is written straight into the packed raster record, RGB and depth share one
4-channel blend pass, and blend-backward's packed gradients are consumed directly by
projection-backward and SH-backward (one kernel, ts_project_sh_bwd).  Forward = 8 launches (projection +
pack + count, SH on a side stream, scan, tile order on a side stream, emit, sort, blend; + the zero fill
of the gradient buffer on the side stream), backward = 2.

The single device->host read (3 ints: intersection count, max list length, oversize tiles) no
longer stalls the step: emit / sort / blend are launched with buffers sized from earlier steps and
the read happens after the blend kernel has been queued (tinysplat_b200/binning.py).
"""
from __future__ import annotations

import os
import weakref
from typing import Tuple

import torch
from torch import Tensor
from torch.autograd import Function

from . import _lib
from . import binning as _binning
from . import rasterize as _rz

BLOCK = 16
_side_streams = {}
USE_SIDE_STREAM = True     # SH kernels on a second stream, overlapping binning / projection-backward
# single-GPU backward tail: projection-backward and SH-backward as ONE kernel (ts_project_sh_bwd) instead of
# two kernels on two streams.  TINYSPLAT_B200_FUSED_TAIL=0 selects the two-kernel form (A/B).
FUSED_TAIL = os.environ.get("TINYSPLAT_B200_FUSED_TAIL", "1") != "0"
# the packed-gradient buffer of blend-backward is zeroed during forward on the side stream (A/B: =0)
PREZERO_GRADS = os.environ.get("TINYSPLAT_B200_PREZERO_GRADS", "1") != "0"
last_bins = None           # (tile_offsets, ids_sorted, M) of the most recent fused forward


def _side_stream(dev, which: int = 0) -> "torch.cuda.Stream":
    key = (str(dev), which)
    if key not in _side_streams:
        _side_streams[key] = torch.cuda.Stream(device=dev)
    return _side_streams[key]


class XysSink:
    """Receives d loss / d xy from the fused node's backward and exposes it the way the
    reference reads it: `extras['xys'].grad.norm(dim=-1)`  [REF model_gaussian.py:130-132]."""
    __slots__ = ("ref",)

    def __init__(self):
        self.ref = None

    def attach(self, xys: Tensor) -> None:
        self.ref = weakref.ref(xys)

    def deliver(self, v_xys: Tensor) -> None:
        t = self.ref() if self.ref is not None else None
        if t is not None:
            t.grad = v_xys if t.grad is None else t.grad + v_xys


class _RenderFused(Function):
    @staticmethod
    def forward(ctx, means, log_scales, quats, opac_logits, colors_dc, colors_rest, view_dev,
                fullproj_dev, fx, fy, width, height, sh_degree, bg4, cull_mode, clamp_rgb, sink, dp, cam_row=None):
        _lib.require_cuda(means, log_scales, quats, opac_logits, colors_dc, colors_rest)
        lib = _lib.load()
        dev = means.device
        st = _lib.stream_ptr(dev)
        N = means.shape[0]
        W, H = int(width), int(height)
        tx, ty = -(-W // BLOCK), -(-H // BLOCK)
        T = tx * ty
        K = colors_rest.shape[1] + 1
        f32 = dict(device=dev, dtype=torch.float32)
        i32 = dict(device=dev, dtype=torch.int32)
        means_c, scales_c, quats_c = (_lib.f32c(t.detach()) for t in (means, log_scales, quats))
        logit_c = _lib.f32c(opac_logits.detach()).reshape(-1)
        dc_c, rest_c = _lib.f32c(colors_dc.detach()), _lib.f32c(colors_rest.detach())
        view_c, proj_c, bg_c = _lib.f32c(view_dev), _lib.f32c(fullproj_dev), _lib.f32c(bg4)
        pflags = _lib.PROJ_LOG_SCALES | _lib.PROJ_RAW_QUATS
        sflags = _lib.SH_DIRS_FROM_MEANS | _lib.SH_OFFSET_CLAMP

        cams_all = None
        if dp is not None:
            # every rank needs every view's camera for the shard backward: 32 floats per rank
            if cam_row is None:
                cam_row = torch.cat([view_c[:3].reshape(-1), proj_c.reshape(-1),
                                     torch.tensor([float(fx), float(fy), 0.0, 0.0], **f32)])
            cam_row = _lib.f32c(cam_row)
            if getattr(dp, "peer", False):
                cams_all = cam_row            # pushed to the peers by ts_dp_push in backward
            else:
                # gathered here so the collective is long finished when backward starts
                cams_all = dp.gather_cameras(cam_row)

        xys = torch.empty(N, 2, **f32)
        depths = torch.empty(N, **f32)
        radii = torch.empty(N, **i32)
        recs = torch.empty(N, lib.ts_rec_floats(), **f32)
        counts = torch.empty(T * lib.ts_bin_counter_stride(), **i32)
        # projection + record packing + tile counting in one pass (conics/cov3d/num_tiles_hit are
        # not materialised: nothing downstream of the fused node reads them)
        _lib.call("ts_project_fwd", N, _lib.ptr(means_c), _lib.ptr(scales_c), 1.0, _lib.ptr(quats_c),
                  _lib.ptr(view_c), _lib.ptr(proj_c), float(fx), float(fy), W / 2, H / 2, H, W, tx, ty,
                  0.01, pflags | _lib.PROJ_OPACITY_LOGIT, _lib.ptr(xys), _lib.ptr(depths), _lib.ptr(radii),
                  None, None, None, _lib.ptr(logit_c), int(cull_mode), _lib.ptr(recs), _lib.ptr(counts), st)
        # SH colour is independent of the bins: it runs on a side stream (ordered after projection,
        # which produced `depths` and the geometry half of `recs`), overlapping scan/emit/sort and
        # keeping the GPU busy while the host waits for the 3 ints below.
        mask = torch.empty(N, device=dev, dtype=torch.uint8)
        main = torch.cuda.current_stream(dev)
        side = _side_stream(dev) if USE_SIDE_STREAM else main
        if side is not main:
            side.wait_stream(main)
        with torch.cuda.stream(side):
            _lib.call("ts_sh_fwd", N, int(sh_degree), K, _lib.ptr(means_c), _lib.ptr(view_c), _lib.ptr(dc_c),
                      _lib.ptr(rest_c), recs.data_ptr() + 32, 12, _lib.ptr(depths), _lib.ptr(mask), sflags,
                      side.cuda_stream)
        offsets = torch.empty(T + 1, **i32)
        stats = torch.empty(lib.ts_bin_scan_work_ints(), **i32)
        _lib.call("ts_bin_scan", T, _lib.ptr(counts), _lib.ptr(offsets), _lib.ptr(stats),
                  lib.ts_bin_smem_sort_cap(), st)
        # emit + sort are launched at once with buffers sized from earlier steps; the 3 ints that say
        # whether that was enough are read after the blend kernel has been queued (binning.py)
        bins = _binning.emit_and_sort(N, T, tx, ty, int(cull_mode), depths, radii, recs, offsets, counts, stats, st)
        if side is not main:
            main.wait_stream(side)          # colours must be in `recs` before blending
        rgb = torch.empty(H, W, 3, **f32)
        depth_img = torch.empty(H, W, **f32)
        final_T = torch.empty(H, W, **f32)
        n_contrib = torch.empty(H, W, **i32)

        def blend():
            _lib.call("ts_blend_fwd", 4, H, W, tx, ty, _lib.ptr(offsets), _lib.ptr(bins.ids_sorted), _lib.ptr(recs),
                      _lib.ptr(bg_c), _lib.ptr(rgb), _lib.ptr(depth_img), _lib.ptr(final_T), _lib.ptr(n_contrib),
                      1 if clamp_rgb else 0, bins.cap_arg, _lib.ptr(bins.order), st)
        blend()
        bins.validate(blend)                # exact lists + a second blend if the capacity was too small
        ids_sorted = bins.ids_sorted
        _rz.last_stats.update(num_intersects=bins.M, max_per_tile=bins.max_count, bins_reused=False)
        global last_bins
        last_bins = (offsets, ids_sorted, bins.M)      # inspection hook (parity tests read the tile lists)
        ctx.save_for_backward(means_c, scales_c, quats_c, logit_c, view_c, proj_c, bg_c, radii, recs,
                              offsets, ids_sorted, final_T, n_contrib, mask)
        ctx.tile_order = bins.order
        # blend-backward accumulates into a zeroed [N,12] buffer (48 MB at 1M): zero it NOW on the side
        # stream, idle while the blend kernel runs, instead of at the head of backward's critical path
        ctx.prezeroed = None
        if PREZERO_GRADS and (dp is None or getattr(dp, "peer", False)) and side is not main and \
                any(ctx.needs_input_grad[:6]):
            with torch.cuda.stream(side):
                buf = torch.zeros(N, lib.ts_grad_floats(), **f32)
                ev = torch.cuda.Event()
                ev.record(side)
            ctx.prezeroed = (buf, ev)
        ctx.meta = (N, K, W, H, tx, ty, float(fx), float(fy), int(sh_degree), pflags, sflags,
                    tuple(opac_logits.shape), tuple(colors_dc.shape))
        ctx.sink = sink
        ctx.dp = dp
        ctx.cams_all = cams_all
        ctx.mark_non_differentiable(xys, depths, radii)
        ctx.set_materialize_grads(False)   # unused outputs (depth, T) arrive as None, not zeros
        # final_T is returned as is (alpha = 1 - T is the caller's one-liner if it wants it):
        # no extra kernel for an output the adapter never reads
        return rgb, depth_img, final_T, xys, depths, radii

    @staticmethod
    def backward(ctx, v_rgb, v_depth, v_T, _vx, _vd, _vr):
        (means_c, scales_c, quats_c, logit_c, view_c, proj_c, bg_c, radii, recs, offsets, ids_sorted,
         final_T, n_contrib, mask) = ctx.saved_tensors
        N, K, W, H, tx, ty, fx, fy, deg, pflags, sflags, opac_shape, dc_shape = ctx.meta
        lib = _lib.load()
        dev = means_c.device
        st = _lib.stream_ptr(dev)
        f32 = dict(device=dev, dtype=torch.float32)
        if v_rgb is None and v_depth is None and v_T is None:
            return (None,) * 19
        v_alpha = -v_T if v_T is not None else None     # alpha = 1 - T
        v_rgb = _lib.f32c(v_rgb) if v_rgb is not None else None
        v_depth = _lib.f32c(v_depth) if v_depth is not None else None
        v_alpha = _lib.f32c(v_alpha) if v_alpha is not None else None
        dp = ctx.dp
        n_rows = N
        peer = dp is not None and getattr(dp, "peer", False)
        if dp is not None and not peer:
            Ns = dp.shard_rows(N)
            n_rows = dp.world * Ns                 # padded so that the all-to-all splits evenly
        split = 1
        if dp is not None and not peer:
            # persistent, zero-initialised: blend-backward clears rows [0, N), the pad stays zero
            grads = dp.buffer("send", (n_rows, lib.ts_grad_floats()), torch.float32, dev, zero=True, tag=N)
        elif getattr(ctx, "prezeroed", None) is not None:
            grads, zeroed = ctx.prezeroed          # zeroed during forward on the side stream
            ctx.prezeroed = None                   # (a second backward of a retained graph allocates its own)
            torch.cuda.current_stream(dev).wait_event(zeroed)
            grads.record_stream(torch.cuda.current_stream(dev))
            split |= _lib.BLEND_GRADS_ZEROED
        else:
            grads = torch.empty(n_rows, lib.ts_grad_floats(), **f32)
        _lib.call("ts_blend_bwd", N, 4, H, W, tx, ty, _lib.ptr(offsets), _lib.ptr(ids_sorted), _lib.ptr(recs),
                  _lib.ptr(bg_c), _lib.ptr(final_T), _lib.ptr(n_contrib), _lib.ptr(v_rgb), _lib.ptr(v_depth), split,
                  _lib.ptr(v_alpha), _lib.ptr(grads), _lib.ptr(ctx.tile_order), st)
        if peer:
            return _RenderFused._backward_peer_exchange(ctx, grads, st)
        if dp is not None:
            return _RenderFused._backward_packed_exchange(ctx, grads, Ns, st)
        # colours first: the largest gradient (colors_rest) becomes available for its all-reduce
        # while projection-backward is still running (parallel.py)
        # all six parameter gradients are views of ONE flat buffer: autograd hands the views to
        # .grad without copying, and the data-parallel reducer (parallel.py) then needs a single
        # NCCL all-reduce over the flat span instead of six
        sizes = [(K - 1) * 3 * N, 3 * N, 3 * N, 3 * N, 4 * N, N]      # rest, dc, means, scales, quats, logit
        offs, total = [], 0
        for sz in sizes:
            offs.append(total)
            total += (sz + 3) // 4 * 4                                # keep every segment 16-byte aligned
        flat = torch.empty(total, **f32)
        seg = lambda i, *shape: flat[offs[i]:offs[i] + sizes[i]].view(*shape)
        v_rest, v_dc = seg(0, N, K - 1, 3), seg(1, N, 1, 3)
        v_means, v_scales, v_quats, v_logit = seg(2, N, 3), seg(3, N, 3), seg(4, N, 4), seg(5, N)
        v_xys = torch.empty(N, 2, **f32)
        if FUSED_TAIL:
            # K6 + K7 in one launch: the SH rows leave as TMA bulk stores while the CTA runs the EWA algebra
            _lib.call("ts_project_sh_bwd", N, deg, K, _lib.ptr(means_c), _lib.ptr(scales_c), 1.0, _lib.ptr(quats_c),
                      _lib.ptr(view_c), _lib.ptr(proj_c), fx, fy, W / 2, H / 2, H, W, pflags | _lib.PROJ_DEPTH_CH3,
                      _lib.ptr(radii), _lib.ptr(grads), _lib.ptr(logit_c), _lib.ptr(mask), _lib.ptr(v_means),
                      _lib.ptr(v_scales), _lib.ptr(v_quats), _lib.ptr(v_logit), _lib.ptr(v_xys), _lib.ptr(v_dc),
                      _lib.ptr(v_rest), st)
        else:
            # SH-backward (DRAM-bound) on the side stream, concurrent with projection-backward
            # (issue-bound); both only read the packed gradients
            main = torch.cuda.current_stream(dev)
            side = _side_stream(dev) if USE_SIDE_STREAM else main
            if side is not main:
                side.wait_stream(main)
            with torch.cuda.stream(side):
                _lib.call("ts_sh_bwd", N, deg, K, _lib.ptr(means_c), _lib.ptr(view_c), grads.data_ptr() + 32, 12,
                          _lib.ptr(mask), _lib.ptr(v_dc), _lib.ptr(v_rest), sflags, side.cuda_stream)
            _lib.call("ts_project_bwd", N, _lib.ptr(means_c), _lib.ptr(scales_c), 1.0, _lib.ptr(quats_c),
                      _lib.ptr(view_c), _lib.ptr(proj_c), fx, fy, W / 2, H / 2, H, W,
                      pflags | _lib.PROJ_DEPTH_CH3, _lib.ptr(radii), None, None, None, _lib.ptr(grads),
                      _lib.ptr(logit_c), _lib.ptr(v_means), _lib.ptr(v_scales), _lib.ptr(v_quats),
                      _lib.ptr(v_logit), _lib.ptr(v_xys), st)
            if side is not main:
                main.wait_stream(side)
        if ctx.sink is not None:
            ctx.sink.deliver(v_xys)
        return (v_means, v_scales, v_quats, v_logit.reshape(opac_shape), v_dc.reshape(dc_shape), v_rest,
                None, None, None, None, None, None, None, None, None, None, None, None, None)

    @staticmethod
    def _backward_packed_exchange(ctx, grads, Ns, st):
        """Data-parallel backward (parallel.PackedGradExchange): this rank's packed rows
        go to the ranks that own the Gaussians (all-to-all, 48 B per Gaussian), the shard backward
        runs for all views at once, the finished shard gradients are all-gathered.  Returns the
        view-summed (or averaged) gradients, identical on every rank."""
        (means_c, scales_c, quats_c, logit_c, view_c, proj_c, bg_c, radii, recs, offsets, ids_sorted,
         final_T, n_contrib, mask) = ctx.saved_tensors
        N, K, W, H, tx, ty, fx, fy, deg, pflags, sflags, opac_shape, dc_shape = ctx.meta
        dp, cams_all = ctx.dp, ctx.cams_all
        dev = means_c.device
        f32 = dict(device=dev, dtype=torch.float32)
        v_xys = torch.empty(N, 2, **f32)
        _lib.call("ts_dp_prepare", N, _lib.ptr(radii), _lib.ptr(mask), _lib.ptr(recs), _lib.ptr(grads),
                  _lib.ptr(v_xys), st)
        recv = dp.all_to_all_rows(grads)                       # [world, Ns, 12]
        s0 = dp.rank * Ns
        ns = max(0, min(N, s0 + Ns) - s0)
        # persistent shard buffers (see PackedGradExchange.buffer), zeroed once: the kernels write
        # rows [0, ns), rows past N stay zero
        shapes = [(Ns, K - 1, 3), (Ns, 1, 3), (Ns, 3), (Ns, 3), (Ns, 4), (Ns,)]
        sh_rest, sh_dc, sh_means, sh_scales, sh_quats, sh_logit = (
            dp.buffer(("shard", i), shp, torch.float32, dev, zero=True, tag=ns) for i, shp in enumerate(shapes))
        scale = dp.out_scale()
        if ns > 0:
            stride = Ns * 12
            main = torch.cuda.current_stream(dev)
            side = _side_stream(dev) if USE_SIDE_STREAM else main
            if side is not main:
                side.wait_stream(main)
            with torch.cuda.stream(side):
                _lib.call("ts_sh_bwd_views", dp.world, ns, deg, K, _lib.ptr(means_c[s0:s0 + ns]),
                          _lib.ptr(cams_all), _lib.ptr(recv), stride, scale, _lib.ptr(sh_dc), _lib.ptr(sh_rest),
                          side.cuda_stream)
            _lib.call("ts_project_bwd_views", dp.world, ns, _lib.ptr(means_c[s0:s0 + ns]),
                      _lib.ptr(scales_c[s0:s0 + ns]), 1.0, _lib.ptr(quats_c[s0:s0 + ns]), _lib.ptr(cams_all),
                      H, W, pflags | _lib.PROJ_DEPTH_CH3, _lib.ptr(recv), stride, _lib.ptr(logit_c[s0:s0 + ns]),
                      scale, _lib.ptr(sh_means), _lib.ptr(sh_scales), _lib.ptr(sh_quats), _lib.ptr(sh_logit), st)
            if side is not main:
                main.wait_stream(side)
        full = dp.all_gather_shards([sh_rest, sh_dc, sh_means, sh_scales, sh_quats, sh_logit])
        # the gathered buffers are reused by the next step: hand autograd its own copies
        v_rest, v_dc, v_means, v_scales, v_quats, v_logit = (t[:N].clone() for t in full)
        if ctx.sink is not None:
            ctx.sink.deliver(v_xys)
        return (v_means, v_scales, v_quats, v_logit.reshape(opac_shape), v_dc.reshape(dc_shape), v_rest,
                None, None, None, None, None, None, None, None, None, None, None, None, None)


def _backward_peer_exchange(ctx, grads, st):
    """Data-parallel backward over NVLink peer memory (parallel.PeerGradExchange,
    csrc/peer.cu): no collective call.  The rows are handled in pieces (PeerLayout.chunks), each
    sharded over the ranks on its own, as a three-stream pipeline:
      main   : ts_dp_push(piece c) — this view's geometry rows to the ranks that own them, its colour
               cotangents + camera to every rank — then SIGNAL(slot c); never waits for a peer;
      side   : WAIT(slot c): every rank's rows of piece c have landed here -> SH gradient of the
               piece's Gaussians rebuilt from all views' colour cotangents (local, HBM-bound);
      side 2 : projection-backward over all views for MY shard of piece c, stored straight into every
               rank's gradient segment (issue-bound + NVLink stores).
    The NVLink transfer of piece c+1 runs under the shard backward of piece c.  A final barrier: every
    shard's gradients have landed in my segment, and the six gradients are views of it."""
    (means_c, scales_c, quats_c, logit_c, view_c, proj_c, bg_c, radii, recs, offsets, ids_sorted,
     final_T, n_contrib, mask) = ctx.saved_tensors
    N, K, W, H, tx, ty, fx, fy, deg, pflags, sflags, opac_shape, dc_shape = ctx.meta
    dp, cam_row = ctx.dp, ctx.cams_all
    dev = means_c.device
    world, rank = dp.world, dp.rank
    L = dp.ensure(N, K, dev)
    _, _, Ns = L.shard_of(rank, N)
    v_xys = torch.empty(N, 2, device=dev, dtype=torch.float32)
    main = torch.cuda.current_stream(dev)
    side, side2 = (_side_stream(dev), _side_stream(dev, 1)) if USE_SIDE_STREAM else (main, main)
    plan, n_pieces, sent = dp.piece_plan(N)
    # one native call issues the whole pipeline (ts_dp_exchange_peer, csrc/peer.cu): ~30 launches, event
    # records and stream waits that would take longer to issue from Python than blend-backward runs
    _lib.call("ts_dp_exchange_peer", N, K, deg, world, rank, n_pieces, plan, world * Ns,
              _lib.ptr(radii), _lib.ptr(mask), _lib.ptr(recs), _lib.ptr(grads), _lib.ptr(cam_row),
              _lib.ptr(means_c), _lib.ptr(scales_c), _lib.ptr(quats_c), _lib.ptr(logit_c),
              dp.bases_table(), dp.seg_offsets_table(), H, W, pflags, dp.out_scale(), dp.next_epoch(),
              float(dp.timeout_s), _lib.ptr(v_xys), main.cuda_stream, side.cuda_stream, side2.cuda_stream)
    # (the main stream has joined the side streams when the call returns: no record_stream needed)
    dp.last_bytes_sent = sent        # bytes this rank sent over NVLink: geometry rows + colours + finished shard gradients
    v_rest, v_dc = dp.local_view("g_rest", N, K - 1, 3), dp.local_view("g_dc", N, 3)
    v_means, v_scales = dp.local_view("g_means", N, 3), dp.local_view("g_scales", N, 3)
    v_quats, v_logit = dp.local_view("g_quats", N, 4), dp.local_view("g_logit", N)
    if ctx.sink is not None:
        ctx.sink.deliver(v_xys)
    return (v_means, v_scales, v_quats, v_logit.reshape(opac_shape), v_dc.reshape(dc_shape), v_rest,
            None, None, None, None, None, None, None, None, None, None, None, None, None)


_RenderFused._backward_peer_exchange = staticmethod(_backward_peer_exchange)


def render_fused(means: Tensor, log_scales: Tensor, quats: Tensor, opacity_logits: Tensor,
                 colors_dc: Tensor, colors_rest: Tensor, view_matrix: Tensor, full_proj: Tensor,
                 fx: float, fy: float, width: int, height: int, sh_degree: int, background: Tensor,
                 cull_mode: int = 1, clamp_rgb: bool = True, grad_exchange=None, cam_row=None
                 ) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor, Tensor]:
    """Fused equivalent of project -> SH(+0.5, clamp) -> rasterise RGB -> rasterise depth.

    Inputs are the raw GaussianModel parameters [REF model_gaussian.py:84-89] and DEVICE camera
    matrices (view 4x4, full projection 4x4).  Returns (rgb[H,W,3] (clamped to <= 1 like the
    adapter's output [REF rasterize.py:45] unless clamp_rgb=False), depth[H,W],
    final_T[H,W] (alpha = 1 - final_T), xys[N,2], depths[N], radii[N]).  `xys.grad` is populated
    by backward.  grad_exchange: a tinysplat_b200.parallel.PackedGradExchange — backward then
    exchanges packed gradient rows between the data-parallel ranks and returns the gradients
    already reduced over all ranks' views (every rank must call with the same N and image size);
    or a parallel.PeerGradExchange — the same over NVLink peer memory without collective calls.
    cam_row: optional device tensor of 32 floats (3x4 view | 4x4 full projection | fx fy 0 0) for
    the exchange; built on the device when absent."""
    sink = XysSink()
    bg = background.to(means.device).float()
    bg4 = torch.cat([bg, bg[:1]])       # depth is composited over background[0] [REF rasterize.py:48-51]
    out = _RenderFused.apply(means, log_scales, quats, opacity_logits, colors_dc, colors_rest,
                             view_matrix, full_proj, fx, fy, width, height, sh_degree, bg4,
                             cull_mode, clamp_rgb, sink, grad_exchange, cam_row)
    sink.attach(out[3])
    return out
