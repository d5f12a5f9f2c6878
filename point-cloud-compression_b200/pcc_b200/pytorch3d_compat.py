"""PyTorch3D-signature front ends of the B200 ops: the names the reference imports at module import time.

  pytorch3d.ops.knn.{_KNN, knn_points, knn_gather}   /root/reference/pn_kit.py:10, train.py:8, compress.py:8,
                                                     eval.py:14, pppe_pcd_ae.py:4
  pytorch3d.ops.{sample_farthest_points, knn_points, knn_gather, ball_query}
                                                     /root/reference/pointnet_sa_module.py:4
  pytorch3d.loss.chamfer_distance                    /root/reference/AE.py:7, PPPF_AE.py:6, eval.py:15,
                                                     pppe_pcd_ae.py:5

Only what the reference exercises is implemented (dense equal-length clouds, norm=2, no normals/weights,
mean/mean reductions); anything else raises NotImplementedError instead of silently falling back.
"""
import torch

from . import ops
from .ops import _KNN


def knn_points(p1, p2, lengths1=None, lengths2=None, norm=2, K=1, version=-1, return_nn=False, return_sorted=True):
    if lengths1 is not None or lengths2 is not None:
        raise NotImplementedError("pcc_b200.knn_points: ragged batches (lengths1/lengths2) are not supported")
    if norm != 2:
        raise NotImplementedError("pcc_b200.knn_points: only norm=2 is supported")
    if p1.shape[0] != p2.shape[0]:
        raise ValueError("pts1 and pts2 must have the same batch dimension.")
    if p1.shape[2] != p2.shape[2]:
        raise ValueError("pts1 and pts2 must have the same point dimension.")
    need_grad_nn = return_nn and p2.requires_grad and torch.is_grad_enabled()
    d, i, nn = ops.knn(p1.detach(), p2.detach(), K, return_nn=return_nn and not need_grad_nn)
    if need_grad_nn:
        nn = ops.gather(p2, i)
    return _KNN(dists=d, idx=i, knn=nn)


def knn_gather(x, idx, lengths=None):
    if lengths is not None:
        raise NotImplementedError("pcc_b200.knn_gather: lengths is not supported")
    if x.shape[0] != idx.shape[0]:
        raise ValueError("x and idx must have same batch dimension")
    return ops.gather(x, idx)


def ball_query(p1, p2, lengths1=None, lengths2=None, K=500, radius=0.2, return_nn=True):
    if lengths1 is not None or lengths2 is not None:
        raise NotImplementedError("pcc_b200.ball_query: ragged batches are not supported")
    d, i = ops.ball_query(p1.detach(), p2.detach(), K, radius)
    nn = None
    if return_nn:  # masked_gather: padded (-1) rows are zero
        nn = ops.gather(p2, i.clamp(min=0)) * (i >= 0).unsqueeze(-1).to(p2.dtype)
    return _KNN(dists=d, idx=i, knn=nn)


def sample_farthest_points(points, lengths=None, K=50, random_start_point=False):
    if lengths is not None or random_start_point:
        raise NotImplementedError("pcc_b200.sample_farthest_points: lengths / random_start_point are not supported")
    if not isinstance(K, int):
        raise NotImplementedError("pcc_b200.sample_farthest_points: per-cloud K is not supported")
    idx = ops.fps(points.detach(), K, None, ops.FLT_MAX)
    pts = ops.gather(points, idx.clamp(min=0))
    if K > points.shape[1]:
        pts = pts * (idx >= 0).unsqueeze(-1).to(pts.dtype)
    return pts, idx


def chamfer_distance(x, y, x_lengths=None, y_lengths=None, x_normals=None, y_normals=None, weights=None,
                     batch_reduction="mean", point_reduction="mean", norm=2, single_directional=False,
                     abs_cosine=True):
    if any(a is not None for a in (x_lengths, y_lengths, x_normals, y_normals, weights)):
        raise NotImplementedError("pcc_b200.chamfer_distance: lengths / normals / weights are not supported")
    if batch_reduction != "mean" or point_reduction != "mean" or norm != 2 or single_directional:
        raise NotImplementedError("pcc_b200.chamfer_distance: only mean/mean, norm=2, bidirectional is supported")
    return ops.chamfer(x, y), None
