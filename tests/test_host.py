"""Host-side logic that needs no GPU: the pytorch3d.* module names, signatures, and loud failure on CPU tensors."""
import inspect
import sys

import pytest
import torch


@pytest.fixture(scope="module")
def pcc():
    import __graft_entry__
    __graft_entry__.build()
    import pcc_b200
    return pcc_b200


def test_install_registers_the_names_the_reference_imports(pcc):
    saved = {k: sys.modules.get(k) for k in ("pytorch3d", "pytorch3d.ops", "pytorch3d.ops.knn", "pytorch3d.loss")}
    try:
        pcc.install()
        from pytorch3d.loss import chamfer_distance  # AE.py:7
        from pytorch3d.ops import ball_query, knn_gather, knn_points, sample_farthest_points  # pointnet_sa_module.py:4
        from pytorch3d.ops.knn import _KNN, knn_gather as kg2, knn_points as kp2  # pn_kit.py:10
        assert knn_points is kp2 and knn_gather is kg2
        assert _KNN._fields == ("dists", "idx", "knn")
        assert chamfer_distance is pcc.chamfer_distance and ball_query is pcc.ball_query
        assert sample_farthest_points is pcc.sample_farthest_points
    finally:
        pcc.uninstall()
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_signatures_follow_pytorch3d(pcc):
    assert list(inspect.signature(pcc.knn_points).parameters) == [
        "p1", "p2", "lengths1", "lengths2", "norm", "K", "version", "return_nn", "return_sorted"]
    assert list(inspect.signature(pcc.ball_query).parameters) == [
        "p1", "p2", "lengths1", "lengths2", "K", "radius", "return_nn"]
    assert list(inspect.signature(pcc.sample_farthest_points).parameters) == [
        "points", "lengths", "K", "random_start_point"]
    assert list(inspect.signature(pcc.chamfer_distance).parameters)[:2] == ["x", "y"]
    assert list(inspect.signature(pcc.farthest_point_sample_batch).parameters) == ["xyz", "npoint"]
    assert list(inspect.signature(pcc.index_points).parameters) == ["points", "idx"]
    for name in ("furthest_point_sample", "ball_query", "group_points", "knn_point"):
        assert hasattr(pcc.PointnetPPOps, name)


def test_cpu_tensors_fail_loudly_no_fallback(pcc):
    x = torch.rand(1, 16, 3)
    with pytest.raises(RuntimeError, match="no CPU path"):
        pcc.knn_points(x, x, K=4)
    with pytest.raises(RuntimeError, match="no CPU path"):
        pcc.farthest_point_sample_batch(x, 4)
    with pytest.raises(RuntimeError, match="no CPU path"):
        pcc.chamfer_distance(x, x)


def test_unsupported_options_raise(pcc):
    x = torch.rand(1, 16, 3)
    with pytest.raises(NotImplementedError):
        pcc.knn_points(x, x, lengths1=torch.tensor([16]), K=4)
    with pytest.raises(NotImplementedError):
        pcc.chamfer_distance(x, x, point_reduction="sum")
    with pytest.raises(ValueError):
        pcc.knn_points(x, torch.rand(2, 16, 3), K=4)


def test_product_never_imports_the_oracle():
    import os
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "point-cloud-compression_b200")
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("CPU oracle", "").replace("the oracle", ""), f


def test_install_swaps_forwards_and_hands_cpu_calls_back(pcc):
    """install(): reference-named modules imported afterwards get their network forwards swapped (import hook); a call that
    cannot run fused (CPU tensor here; training / train-mode BatchNorm on the GPU box) runs the class's own forward."""
    import importlib
    import os
    names = ("pn_kit", "AE", "pointnet_sa_module", "PPPF_AE")
    shims = ("pytorch3d", "pytorch3d.ops", "pytorch3d.ops.knn", "pytorch3d.loss", "torchac")
    saved = {k: sys.modules.pop(k, None) for k in names + shims}
    standins = os.path.join(os.path.dirname(os.path.abspath(__file__)), "standins")
    sys.path.insert(0, standins)
    try:
        pcc.install()
        ae_mod = importlib.import_module("AE")                       # pulls pn_kit in; both are patched when their import ends
        pn = sys.modules["pn_kit"]
        assert pn.MLP.forward.__pcc_b200__ and pn.SetAbstraction.forward.__pcc_b200__ and ae_mod.AE.forward.__pcc_b200__
        assert pn.farthest_point_sample_batch is pcc.farthest_point_sample_batch
        mlp = pn.MLP(in_channel=8, mlps=[16, 4], relu=[True, False], bn=False).eval()
        x = torch.rand(2, 8, 10)
        with torch.no_grad():
            y = mlp(x)                                               # CPU tensor: the class's own forward
            want = pn.MLP.forward.__pcc_original__(mlp, x)
        assert y.shape == (2, 4, 10) and torch.equal(y, want)
        pcc.uninstall()
        assert not hasattr(pn.MLP.forward, "__pcc_b200__")
    finally:
        pcc.uninstall()
        sys.path.remove(standins)
        for k, v in saved.items():
            sys.modules.pop(k, None)
            if v is not None:
                sys.modules[k] = v
