"""knn_points at the density regularizer's size: 100k sampled points vs 1M Gaussian means, K=16."""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from tinysplat_b200.knn import knn_points
P1, P2, K = 100_000, 1_000_000, 16
g = torch.Generator().manual_seed(0)
p1 = torch.randn(1, P1, 3, generator=g).cuda(); p2 = torch.randn(1, P2, 3, generator=g).cuda()
knn_points(p1, p2, K=K); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); out = knn_points(p1, p2, K=K); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
# spot-check 64 queries against torch
d = ((p1[0, :64, None, :] - p2[0][None]) ** 2).sum(-1)
wd, wi = torch.topk(d, K, dim=1, largest=False)
print(json.dumps({"queries": P1, "refs": P2, "K": K, "ms": round(ms, 2), "pairs_per_s": round(P1 * P2 / ms * 1e3 / 1e12, 2),
                  "unit": "T pair-distances/s", "spot_check_idx_match": float((out.idx[0, :64] == wi).float().mean())}))
