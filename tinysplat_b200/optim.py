"""Fused Adam for the Gaussian parameters.

Drop-in for the optimizer the reference builds, `optim.Adam(model.parameters())` over six named
parameter groups [REF scripts/train.py:26; tinysplat/splatting/model_gaussian.py:112-120]:

    optimizer = tinysplat_b200.optim.FusedAdam(model.parameters())

It subclasses torch.optim.Adam and keeps torch's state layout (`state[p]['step']`,
`['exp_avg']`, `['exp_avg_sq']`), so the reference's densify/prune optimizer surgery — which
masks and concatenates exactly those tensors and re-keys the state under a new Parameter
[REF model_gaussian.py:199-242] — keeps working unchanged.  step() updates every tensor in ONE
kernel launch (ts_adam_step); there is no CPU path."""
from __future__ import annotations

import ctypes as C

import torch
from torch.optim import Adam

from . import _lib


class FusedAdam(Adam):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=False)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        max_t = lib.ts_adam_max_tensors()
        for group in self.param_groups:
            if group.get("weight_decay", 0) != 0 or group.get("amsgrad", False) or group.get("maximize", False):
                raise NotImplementedError("FusedAdam supports plain Adam only (no weight decay / amsgrad / maximize)")
        # batch tensors that share (beta1, beta2, eps) into launches of up to max_t tensors
        batches = {}
        for group in self.param_groups:
            b1, b2 = group["betas"]
            key = (float(b1), float(b2), float(group["eps"]))
            for p in group["params"]:
                if p.grad is None:
                    continue
                _lib.require_cuda(p)
                if p.grad.is_sparse:
                    raise NotImplementedError("FusedAdam does not support sparse gradients")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                if not (p.is_contiguous() and st["exp_avg"].is_contiguous() and st["exp_avg_sq"].is_contiguous()
                        and p.dtype == torch.float32):
                    raise NotImplementedError("FusedAdam needs contiguous fp32 parameters and state")
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                batches.setdefault(key, []).append((p, g, st, float(group["lr"])))
        for (b1, b2, eps), items in batches.items():
            for i in range(0, len(items), max_t):
                chunk = items[i:i + max_t]
                n = len(chunk)
                arr = lambda vals: (C.c_void_p * n)(*vals)
                _lib.call("ts_adam_step", n,
                          arr([p.data_ptr() for p, _, _, _ in chunk]),
                          arr([g.data_ptr() for _, g, _, _ in chunk]),
                          arr([s["exp_avg"].data_ptr() for _, _, s, _ in chunk]),
                          arr([s["exp_avg_sq"].data_ptr() for _, _, s, _ in chunk]),
                          (C.c_int64 * n)(*[p.numel() for p, _, _, _ in chunk]),
                          (C.c_float * n)(*[lr for _, _, _, lr in chunk]),
                          (C.c_int64 * n)(*[int(s["step"].item()) for _, _, s, _ in chunk]),
                          b1, b2, eps, _lib.stream_ptr(chunk[0][0].device))
                # the kernel wrote through raw pointers: tell autograd (and every cache keyed on
                # Tensor._version, e.g. the tile-bin cache) that these tensors changed
                for p, _, s, _ in chunk:
                    torch.autograd.graph.increment_version(p)
                    torch.autograd.graph.increment_version(s["exp_avg"])
                    torch.autograd.graph.increment_version(s["exp_avg_sq"])
        return loss
