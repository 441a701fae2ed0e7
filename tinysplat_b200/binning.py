"""Host side of K3 after the tile scan: emit + per-tile sort WITHOUT waiting for the intersection
count  [behind gsplat.rasterize_gaussians, REF tinysplat/splatting/rasterize.py:44,50].

The number of (tile, Gaussian) pairs M is only known on the device after the scan, and the key / id
buffers have to be sized on the host.  Round 1 read M back and stalled the host in the middle of every
forward pass.  Now the buffers are sized from what earlier calls needed (+50 %), emit / sort / blend
are launched immediately with that capacity — the kernels never write or read past it — and the three
integers (M, longest list, oversized tiles) are read AFTER the blend kernel has been queued: by then
the copy has long finished, so the host does not drain the GPU, and the GPU always has a few hundred
microseconds of queued work while the host runs ahead.  If the step did need more than the capacity
(first call, scene change), the pass is repeated with exact sizes: results are always exact.
"""
from __future__ import annotations

import os
from typing import Callable, Dict, Optional, Tuple

import torch
from torch import Tensor

from . import _lib

# (device index, tiles) -> [capacity, max-list bound, oversized tiles last seen]; never shrinks (memory only)
_state: Dict[Tuple[int, int], list] = {}
_pinned: Dict[int, list] = {}
stats = {"speculative": 0, "exact_first": 0, "redone": 0}      # counters (tests, bench run_info)

# A/B switch (bench only): TINYSPLAT_B200_TILE_ORDER=0 launches the blend kernels in raster order
USE_TILE_ORDER = os.environ.get("TINYSPLAT_B200_TILE_ORDER", "1") != "0"
HEADROOM = 1.5          # key / id buffers: memory only (12 B per pair), a too-small guess costs a repeated pass
LIST_HEADROOM = 1.25    # longest-list bound: decides which sort size classes are launched, so kept tight
SLACK = 4096


def reset_state() -> None:
    _state.clear()


def _host_buffer(dev) -> Tensor:
    """A 4-int pinned landing slot; a small ring so that several binnings can be pending at once."""
    k = dev.index if dev.index is not None else torch.cuda.current_device()
    if k not in _pinned:
        _pinned[k] = [torch.empty(16, 4, dtype=torch.int32).pin_memory(), 0]
    ring = _pinned[k]
    ring[1] = (ring[1] + 1) % 16
    return ring[0][ring[1]]


class PendingBins:
    """Tile lists of one binning; `validate()` must be called (after the consumer kernel has been
    queued) before anything else trusts M — it reads the counts back and, if the speculative
    capacity was too small, rebuilds the lists exactly and asks the caller to run its consumer again."""
    __slots__ = ("offsets", "ids_sorted", "keys", "order", "M", "max_count", "capacity", "cap_arg", "_ev", "_redo",
                 "_host", "_key")

    def __init__(self):
        self.M = None
        self.max_count = None

    def validate(self, rerun_consumer: Optional[Callable[[], None]] = None) -> bool:
        """Returns True when the lists had to be rebuilt (rerun_consumer, if given, has been called)."""
        if self._ev is None:
            return False
        self._ev.synchronize()
        self._ev = None
        M, max_count, n_big, _ = self._host.tolist()
        self.M, self.max_count = M, max_count
        cap_used, bound_used = self.capacity
        st = _state.setdefault(self._key, [0, 0, 0])
        st[0] = max(st[0], int(M * HEADROOM) + SLACK)
        st[1] = max(st[1], int(max_count * LIST_HEADROOM) + 1)
        st[2] = n_big
        if M <= cap_used and max_count <= bound_used and n_big == 0:
            return False
        stats["redone"] += 1
        self._redo(self, M, max_count, n_big)
        if rerun_consumer is not None:
            rerun_consumer()
        return True


_order_streams: Dict[int, "torch.cuda.Stream"] = {}


def emit_and_sort(N: int, T: int, tx: int, ty: int, cull_mode: int, depths: Tensor, radii: Tensor,
                  recs: Tensor, offsets: Tensor, counts: Tensor, scan_stats: Tensor, stream_ptr: int,
                  order_stream=None) -> PendingBins:
    """Launches ts_bin_emit + ts_bin_sort for the scanned tile counts.  `counts` holds the emit cursors
    (ts_bin_scan turned the counters into cursors), `scan_stats` the scan's device-side statistics.
    The blend kernels' launch order (ts_bin_tile_order, one small CTA) runs beside emit / sort on
    `order_stream` (default: a stream of this module); the caller's stream is made to wait for it."""
    lib = _lib.load()
    dev = depths.device
    main = torch.cuda.current_stream(dev)
    host = _host_buffer(dev)
    with torch.cuda.stream(main):
        host.copy_(scan_stats[:4], non_blocking=True)
    ev = torch.cuda.Event()
    ev.record(main)
    key = (dev.index if dev.index is not None else torch.cuda.current_device(), T)
    bins = PendingBins()
    bins.offsets, bins._host, bins._key = offsets, host, key
    i64 = dict(device=dev, dtype=torch.int64)
    i32 = dict(device=dev, dtype=torch.int32)

    def launch(b: PendingBins, cap: int, max_count: int, n_big: int, exact: bool):
        b.keys = torch.empty(max(cap, 1), **i64)
        b.ids_sorted = torch.empty(max(cap, 1), **i32)
        if N == 0 or cap <= 0:
            return
        _lib.call("ts_bin_emit", N, _lib.ptr(depths), _lib.ptr(radii), _lib.ptr(recs), tx, ty, int(cull_mode),
                  _lib.ptr(offsets), _lib.ptr(counts), _lib.ptr(b.keys), 0 if exact else cap, stream_ptr)
        big_scratch = big_counter = None
        if n_big > 0:
            P = 1 << (max_count - 1).bit_length()
            big_scratch = torch.empty(n_big * P, **i64)
            big_counter = torch.empty(1, **i32)
        _lib.call("ts_bin_sort", T, _lib.ptr(offsets), _lib.ptr(b.keys), _lib.ptr(b.ids_sorted), max_count, n_big,
                  _lib.ptr(big_scratch), _lib.ptr(big_counter), 0 if exact else cap, stream_ptr)

    def redo(b: PendingBins, M: int, max_count: int, n_big: int):
        # emit advanced the cursors: restore them from the offsets, then rebuild with exact sizes
        _lib.call("ts_bin_reset_cursors", T, _lib.ptr(offsets), _lib.ptr(counts), stream_ptr)
        launch(b, M, max_count, n_big, True)
        b.capacity, b.cap_arg = (M, max_count), 0

    bins._redo = redo
    # launch order of the blend kernels (longest lists first); depends on the offsets only
    bins.order = None
    order_done = None
    if USE_TILE_ORDER:
        bins.order = torch.empty(T, **i32)
        if order_stream is None:
            order_stream = _order_streams.get(key[0])
            if order_stream is None:
                order_stream = _order_streams[key[0]] = torch.cuda.Stream(device=dev)
        order_stream.wait_event(ev)            # the scan has produced the offsets
        _lib.call("ts_bin_tile_order", T, _lib.ptr(offsets), _lib.ptr(bins.order), order_stream.cuda_stream)
        bins.order.record_stream(order_stream)
        order_done = torch.cuda.Event()
        order_done.record(order_stream)
    st = _state.get(key)
    if st is None or st[2] > 0:
        # first binning of this (device, image size): nothing to extrapolate from — size exactly
        # (also while some tile's list exceeds the shared-memory sort: its scratch is sized from n_big)
        ev.synchronize()
        M, max_count, n_big, _ = host.tolist()
        stats["exact_first"] += 1
        bins.M, bins.max_count, bins._ev = M, max_count, None
        bins.capacity, bins.cap_arg = (M, max_count), 0
        launch(bins, M, max_count, n_big, True)
        prev = st or [0, 0, 0]
        _state[key] = [max(prev[0], int(M * HEADROOM) + SLACK), max(prev[1], int(max_count * LIST_HEADROOM) + 1), n_big]
        if order_done is not None:
            main.wait_event(order_done)
        return bins
    stats["speculative"] += 1
    bins._ev = ev
    bins.capacity, bins.cap_arg = (st[0], st[1]), st[0]      # cap_arg: what ts_blend_fwd must be told (0 = exact)
    launch(bins, st[0], st[1], 0, False)
    if order_done is not None:
        main.wait_event(order_done)
    return bins
