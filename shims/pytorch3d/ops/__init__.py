"""`from pytorch3d.ops import knn_points, ball_query`  [REF tinysplat/splatting/model_gaussian.py:16].

`knn_points` is tinysplat_b200's exact brute-force sm_100a kernel (tinysplat_b200/knn.py) with
pytorch3d's signature and return fields, as the reference calls it
(`knn_points(points[None], means[None], K=16).idx[0]`  [REF model_gaussian.py:260,425,519]).
`ball_query` is imported by the reference but never called on the training path; it raises."""
from tinysplat_b200.knn import knn_points  # noqa: F401


def ball_query(*args, **kwargs):
    raise NotImplementedError("pytorch3d.ops.ball_query is not provided by the tinysplat_b200 shim "
                              "(the reference imports it but does not call it while training)")


__all__ = ["knn_points", "ball_query"]
