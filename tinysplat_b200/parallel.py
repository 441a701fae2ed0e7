"""Data-parallel rendering over cameras (new: the reference renders one camera per step on one
device [REF scripts/train.py:54-55]).

One process per GPU, each holding a full replica of the Gaussian parameters.  Every rank
renders its own view(s) forward+backward; the only exchange step is a sum of the per-Gaussian
parameter gradients (59 floats per Gaussian at SH degree 3).  Each parameter's all-reduce is
launched from a post-accumulate-grad hook the moment autograd has produced that gradient, so
the largest payload (colors_rest, 76% of the bytes, ready right after SH-backward) crosses
NVLink while projection-backward is still running.  Forward-only rendering needs no
collective at all (replicas only).

The densification statistic is the per-view norm of d loss / d xy summed over views
[REF tinysplat/splatting/model_gaussian.py:130-132] — NOT the norm of the reduced gradient —
so it gets its own [N] all-reduce (`reduce_densify_stat`)."""
from __future__ import annotations

from typing import Callable, Iterable, List, Optional

import torch
import torch.distributed as dist
from torch import Tensor


class GradientAllReducer:
    """All-reduce (mean or sum) of .grad for a fixed list of leaf tensors.

    Launch happens from the post-accumulate-grad hook of the LAST parameter to receive its
    gradient, i.e. as early as autograd allows.  Gradients that are views of one flat buffer
    (the fused node allocates them that way) are reduced with a single collective over the
    buffer's span; anything else falls back to one collective per tensor."""

    def __init__(self, params: Iterable[Tensor], process_group=None, average: bool = True,
                 overlap: bool = True):
        self.params: List[Tensor] = list(params)
        self.group = process_group
        self.average = average
        self.overlap = overlap
        self._works = []       # (tensor that was reduced, work handle)
        self._handles = []
        self._ready = 0
        self._launched = False
        self.last_num_collectives = 0
        self._enabled = dist.is_available() and dist.is_initialized() and \
            dist.get_world_size(process_group) > 1
        self._native_avg = self._enabled and dist.get_backend(process_group) == "nccl"
        if self._enabled and overlap:
            for p in self.params:
                self._handles.append(p.register_post_accumulate_grad_hook(self._hook))

    @property
    def world_size(self) -> int:
        return dist.get_world_size(self.group) if self._enabled else 1

    def _hook(self, p: Tensor) -> None:
        self._ready += 1
        if self._ready == len(self.params):
            self._launch()

    def _spans(self):
        """Group gradients by storage; a group whose members tile one span densely becomes one
        flat tensor over that span."""
        groups = {}
        for p in self.params:
            if p.grad is not None:
                groups.setdefault(p.grad.untyped_storage().data_ptr(), []).append(p.grad)
        out = []
        for grads in groups.values():
            if len(grads) > 1 and all(g.is_contiguous() for g in grads):
                lo = min(g.storage_offset() for g in grads)
                hi = max(g.storage_offset() + g.numel() for g in grads)
                used = sum(g.numel() for g in grads)
                if hi - lo <= used + 4 * len(grads):        # only alignment padding in between
                    flat = torch.empty(0, dtype=grads[0].dtype, device=grads[0].device)
                    flat.set_(grads[0].untyped_storage(), lo, (hi - lo,))
                    out.append(flat)
                    continue
            out.extend(grads)
        return out

    def _launch(self) -> None:
        op = dist.ReduceOp.AVG if (self.average and self._native_avg) else dist.ReduceOp.SUM
        for t in self._spans():
            self._works.append((t, dist.all_reduce(t, op=op, group=self.group, async_op=True)))
        self.last_num_collectives = len(self._works)
        self._launched = True

    def finish(self) -> None:
        """Wait for the in-flight reductions (launching them now if the hooks did not)."""
        if not self._enabled:
            return
        if not self._launched:
            self._launch()
        for t, w in self._works:
            w.wait()
            if self.average and not self._native_avg:
                t.div_(self.world_size)
        self._works.clear()
        self._ready = 0
        self._launched = False

    def payload_bytes(self) -> int:
        return sum(p.numel() * p.element_size() for p in self.params)

    def close(self) -> None:
        for h in self._handles:
            h.remove()
        self._handles.clear()


class DataParallelRenderer:
    """rank r renders cameras[r::world] with `rasterizer`, backpropagates `loss_fn`, and leaves
    the view-averaged gradient in every parameter's .grad on every rank."""

    def __init__(self, rasterizer: Callable, params: Iterable[Tensor], process_group=None,
                 average: bool = True, overlap: bool = True):
        self.rasterizer = rasterizer
        self.reducer = GradientAllReducer(params, process_group, average, overlap)
        self.group = process_group

    def step(self, camera, dims, sh_degree: int, loss_fn: Callable[[Tensor, dict], Tensor]):
        img, extras = self.rasterizer(camera, dims, sh_degree)
        loss = loss_fn(img, extras)
        loss.backward()
        self.reducer.finish()
        return loss.detach(), img.detach(), extras

    def render_only(self, camera, dims, sh_degree: int):
        with torch.no_grad():
            return self.rasterizer(camera, dims, sh_degree)

    def reduce_densify_stat(self, extras: dict) -> Optional[Tensor]:
        g = extras["xys"].grad
        if g is None:
            return None
        stat = g.norm(dim=-1)
        if self.reducer._enabled:
            dist.all_reduce(stat, op=dist.ReduceOp.SUM, group=self.group)
        return stat
