"""Exact K-nearest-neighbour search: stand-in for `pytorch3d.ops.knn_points`,
which the reference's density regularizer and mesh export call as
`knn_points(points[None], self.means[None], K=16).idx[0]`
[REF tinysplat/splatting/model_gaussian.py:16,260,425,519]; pytorch3d is not installable here.

    from tinysplat_b200.knn import knn_points

Returns the same namedtuple fields (dists = squared L2 ascending, idx int64, knn).  Indices and
distances carry no gradient (the reference only indexes with them).  No CPU path."""
from __future__ import annotations

from collections import namedtuple
from typing import Optional

import torch
from torch import Tensor

from . import _lib

_KNN = namedtuple("KNN", "dists idx knn")


@torch.no_grad()
def knn_points(p1: Tensor, p2: Tensor, lengths1: Optional[Tensor] = None, lengths2: Optional[Tensor] = None,
               norm: int = 2, K: int = 1, version: int = -1, return_nn: bool = False,
               return_sorted: bool = True) -> _KNN:
    if norm != 2:
        raise NotImplementedError("knn_points: only the L2 norm is supported")
    if lengths1 is not None or lengths2 is not None:
        raise NotImplementedError("knn_points: ragged batches (lengths1/lengths2) are not supported")
    if p1.dim() != 3 or p2.dim() != 3 or p1.shape[0] != p2.shape[0] or p1.shape[2] != 3 or p2.shape[2] != 3:
        raise ValueError("knn_points expects p1 [N,P1,3] and p2 [N,P2,3]")
    if K not in (1, 2, 4, 8, 16, 32):
        raise NotImplementedError("knn_points: K must be one of 1, 2, 4, 8, 16, 32")
    _lib.require_cuda(p1, p2)
    N, P1, _ = p1.shape
    P2 = p2.shape[1]
    dev = p1.device
    dists = torch.empty(N, P1, K, device=dev, dtype=torch.float32)
    idx = torch.empty(N, P1, K, device=dev, dtype=torch.int64)
    for b in range(N):
        q = _lib.f32c(p1[b].detach())
        r = _lib.f32c(p2[b].detach())
        _lib.call("ts_knn_points", P1, P2, K, _lib.ptr(q), _lib.ptr(r), _lib.ptr(dists[b]), _lib.ptr(idx[b]),
                  _lib.stream_ptr(dev))
    nn = None
    if return_nn:
        nn = torch.gather(p2[:, None].expand(N, P1, P2, 3), 2, idx[..., None].expand(N, P1, K, 3))
    return _KNN(dists=dists, idx=idx, knn=nn)
