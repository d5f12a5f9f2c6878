"""Import the reference's own Python modules from /root/reference (THIS container only).

TEST INFRASTRUCTURE ONLY.  /root/reference does not exist on the GPU box, so nothing under `-m gpu`, smoke()
or bench.py may call this; it is used by tests/golden/make_golden.py (to mint golden vectors from the
reference's own code) and by CPU tests that are skipped when the reference tree is absent.

The reference imports third-party packages that are not installed here (pytorch3d, pyntcloud, plyfile,
torchac, open3d).  They are stubbed in sys.modules; the pytorch3d stubs are backed by the CPU oracle
(oracle/oracle.py), i.e. the restated PyTorch3D semantics of SURVEY.md appendix A.
"""
import collections
import importlib
import os
import sys
import types

import numpy as np
import torch

from . import oracle as orc

REFERENCE_ROOT = "/root/reference"
_KNN = collections.namedtuple("KNN", "dists idx knn")


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "pn_kit.py"))


def _t(a, like=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    return t


def _oracle_knn_points(p1, p2, lengths1=None, lengths2=None, norm=2, K=1, version=-1, return_nn=False,
                       return_sorted=True):
    d, i, nn = orc.knn_points(p1.detach().numpy(), p2.detach().numpy(), K, return_nn)
    return _KNN(_t(d), _t(i), _t(nn) if nn is not None else None)


def _oracle_knn_gather(x, idx, lengths=None):
    # differentiable torch statement of knn_gather (appendix A) so reference modules can be back-propagated
    N, L, K = idx.shape
    U = x.shape[2]
    return x[:, :, None].expand(-1, -1, K, -1).gather(1, idx[..., None].expand(-1, -1, -1, U))


def _oracle_ball_query(p1, p2, lengths1=None, lengths2=None, K=500, radius=0.2, return_nn=True):
    d, i = orc.ball_query(p1.detach().numpy(), p2.detach().numpy(), K, radius)
    nn = None
    if return_nn:
        g = orc.gather(p2.detach().numpy(), np.maximum(i, 0))
        g[i < 0] = 0
        nn = _t(g)
    return _KNN(_t(d), _t(i), nn)


def _oracle_sample_farthest_points(points, lengths=None, K=50, random_start_point=False):
    pts, idx = orc.sample_farthest_points(points.detach().numpy(), K)
    return _t(pts), _t(idx)


class _OracleChamfer(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y):
        loss, _, _, ix, _, iy = orc.chamfer(x.detach().numpy(), y.detach().numpy())
        ctx.save_for_backward(x, y, _t(ix), _t(iy))
        return torch.tensor(loss, dtype=torch.float32)

    @staticmethod
    def backward(ctx, g):
        x, y, ix, iy = ctx.saved_tensors
        gx, gy = orc.chamfer_bwd(x.detach().numpy(), y.detach().numpy(), ix.numpy(), iy.numpy(), float(g))
        return _t(gx), _t(gy)


def _oracle_chamfer_distance(x, y, *args, **kwargs):
    return _OracleChamfer.apply(x, y), None


def install_stubs():
    """Put stub modules for the reference's absent third-party imports into sys.modules."""
    def mod(name):
        m = types.ModuleType(name)
        sys.modules[name] = m
        return m

    if "pytorch3d" not in sys.modules or not hasattr(sys.modules["pytorch3d"], "__pcc_oracle__"):
        p3d = mod("pytorch3d")
        p3d.__pcc_oracle__ = True
        ops = mod("pytorch3d.ops")
        knn = mod("pytorch3d.ops.knn")
        loss = mod("pytorch3d.loss")
        p3d.ops, p3d.loss, ops.knn = ops, loss, knn
        for m in (ops, knn):
            m._KNN = _KNN
            m.knn_points = _oracle_knn_points
            m.knn_gather = _oracle_knn_gather
        ops.ball_query = _oracle_ball_query
        ops.sample_farthest_points = _oracle_sample_farthest_points
        loss.chamfer_distance = _oracle_chamfer_distance
    for name, attrs in (("pyntcloud", ["PyntCloud"]), ("plyfile", ["PlyData", "PlyElement"]), ("torchac", []),
                        ("open3d", [])):
        if name not in sys.modules:
            m = mod(name)
            for a in attrs:
                setattr(m, a, type(a, (), {}))


def load(name):
    """Import reference module `name` (pn_kit, AE, PPPF_AE, pointnet_sa_module, pppe_pcd_ae, octree_np)."""
    if not available():
        raise RuntimeError("/root/reference is not present")
    install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    return importlib.import_module(name)
